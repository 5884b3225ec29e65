"""Basic-block breakdown of one kernel of an .ncu-rep (SASS page): instructions executed, share of samples, lanes.
    python tools/ncu_blocks.py gpurun_out/x.ncu-rep [min_share_pct] [kernel-name regex]"""
import csv, subprocess, sys
sel = ["--kernel-name", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv"] + sel, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.4
hdr = rows[1]; data = rows[2:]
iS, iI, iT, iSm = (hdr.index(k) for k in ("Source", "Instructions Executed", "Avg. Threads Executed", "# Samples"))
tot = sum(int(r[iI]) for r in data); tots = sum(int(r[iSm]) for r in data)
print(rows[0][1]); print("total warp-instructions %.3e, samples %d" % (tot, tots))
seg, cur = [], None
for k, r in enumerate(data):
    n = int(r[iI])
    if cur is None or abs(n - cur["n"]) > 0.02 * max(n, cur["n"], 1):
        cur = {"n": n, "start": k, "cnt": 0, "inst": 0, "samp": 0, "thr": 0.0, "text": []}; seg.append(cur)
    cur["cnt"] += 1; cur["inst"] += n; cur["samp"] += int(r[iSm]); cur["thr"] += float(r[iT]); cur["text"].append(r[iS].strip())
for s in seg:
    if s["inst"] > thr / 100 * tot or s["samp"] > thr / 100 * tots:
        ops = " ".join(t.split()[0] if not t.startswith("@") else t.split()[1] for t in s["text"][:12])
        print("%5d len %3d x %8.3fM  inst %5.1f%%  samples %5.1f%%  lanes %4.1f | %s" % (s["start"], s["cnt"], s["n"] / 1e6, 100 * s["inst"] / tot, 100 * s["samp"] / tots, s["thr"] / s["cnt"], ops[:130]))
