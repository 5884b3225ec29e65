"""Per-kernel table from an `ncu --csv --metrics ...` launch list (one row per kernel x metric)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
h = rows[hdr]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    per.setdefault((r[ii], r[ki].split("(")[0][-48:]), {})[r[mi]] = v
for (i, k), m in per.items():
    t = m.get("gpu__time_duration.sum", 0.0)
    rd, wr = m.get("dram__bytes_read.sum", 0.0), m.get("dram__bytes_write.sum", 0.0)
    print(f"{i:>4s} {k:48s} {t / 1e6:8.3f} ms  rd {rd / 1e9:7.3f} GB  wr {wr / 1e9:7.3f} GB  "
          f"L2hit {m.get('lts__t_sector_hit_rate.pct', 0.0):5.1f}%  {(rd + wr) / max(t, 1.0):7.2f} GB/s/1e0" )
