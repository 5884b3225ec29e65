"""Extract per-launch DRAM traffic (dram__bytes_read.sum + dram__bytes_write.sum) of one kernel from an
.ncu-rep into profiles/roofline_traffic.json, which bench.py reads for `roofline.traffic`.

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep dsm_gather_kernel joint_10k
"""
import csv, json, os, subprocess, sys

def main(path, kernel, workload):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    vals = []
    for r in rows[2:]:
        if kernel in r[ki]:
            t = 0.0
            for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                i = hdr.index(m)
                t += float(r[i]) * scale.get(units[i], 1)
            vals.append(t)
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "roofline_traffic.json")
    data = json.load(open(out)) if os.path.exists(out) else {}
    data.setdefault(workload, {})[kernel] = {"traffic_bytes_per_launch": sum(vals) / len(vals), "launches": len(vals),
                                             "source": os.path.basename(path)}
    json.dump(data, open(out, "w"), indent=1, sort_keys=True)
    print(data[workload][kernel])

if __name__ == "__main__":
    main(*sys.argv[1:4])
