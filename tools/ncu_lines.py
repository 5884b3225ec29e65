"""Per-source-line breakdown of one kernel of an .ncu-rep (needs -lineinfo + --import-source on):
    python tools/ncu_lines.py gpurun_out/x.ncu-rep [top_n]"""
import csv, subprocess, sys, os
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(raw.splitlines()))
fname, hdr, out = None, None, []
for r in rows:
    if not r: continue
    if r[0] in ("File Name", "File Path"): fname = os.path.basename(r[1]); hdr = None; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) < len(hdr): continue
    if r[2] != "-": continue   # SASS rows follow their CUDA line's aggregate row (Address "-")
    try:
        ii, isamp, ithr = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Thread Instructions Executed")
        n = int(r[ii]); s = int(r[isamp]); t = int(r[ithr])
    except (ValueError, IndexError):
        continue
    if n or s: out.append((fname, int(r[0]), n, s, t, r[1].strip()))
tot = sum(o[2] for o in out); tots = sum(o[3] for o in out)
print("total warp-instructions %.3e samples %d" % (tot, tots))
for f, ln, n, s, t, src in sorted(out, key=lambda o: -o[2])[:top]:
    print("%-22s %4d  inst %5.1f%%  samp %5.1f%%  lanes %4.1f | %s" % (f[:22], ln, 100 * n / tot, 100 * s / max(tots, 1), t / max(n, 1), src[:90]))
