"""Summarise an .ncu-rep (read here, no GPU needed) into the few numbers DESIGN.md / bench.py cite.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep > profiles/x_summary.txt
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__grid_size", "launch__block_size",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__maximum_warps_per_active_cycle_pct"]


def main(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("# %s" % path)
    for r in rows[2:]:
        print("\n== %s" % r[ki])
        rd = wr = None
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print("  %-72s %s %s" % (w, r[i], units[i]))
                if w == "dram__bytes_read.sum":
                    rd = (float(r[i]), units[i])
                if w == "dram__bytes_write.sum":
                    wr = (float(r[i]), units[i])
        if rd and wr:
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tot = rd[0] * scale.get(rd[1], 1) + wr[0] * scale.get(wr[1], 1)
            print("  %-72s %.0f byte" % ("traffic = dram read + write", tot))


if __name__ == "__main__":
    main(sys.argv[1])
