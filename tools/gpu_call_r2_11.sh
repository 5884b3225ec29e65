#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c11.log; : > $L
run() { echo "== $*" >> $L; timeout 600 "$@" 2>gpurun_out/r2_c11.err | tail -1 | cut -c1-6000 >> $L || echo "FAILED rc=$?" >> $L; grep -v "^$" gpurun_out/r2_c11.err | tail -3 | cut -c1-400 >> $L; }
run python bench.py --steps 5 --warmup 2 --no-cpu-baseline
run env AMB_COMPACT_MIRRORS=1 python bench.py --steps 5 --warmup 2 --no-cpu-baseline
run env AMB_DSM_STREAM_CHUNKS=4 python bench.py --steps 5 --warmup 2 --no-cpu-baseline
run env AMB_COMPACT_MIRRORS=1 AMB_DSM_STREAM_CHUNKS=4 python bench.py --steps 5 --warmup 2 --no-cpu-baseline
run env AMB_COMPACT_MIRRORS=1 AMB_DSM_STREAM_CHUNKS=8 python bench.py --steps 5 --warmup 2 --no-cpu-baseline
timeout 300 python -m pytest -m gpu -q -x tests/test_gpu_compact_mirrors.py 2>&1 | tail -2 >> $L
python - $L <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("=="): print(ln.strip()[:200]); continue
    try: d = json.loads(ln)
    except Exception: print("   ", ln.strip()[:400]); continue
    e = d.get("e2e") or {}
    print("    step %.3f ms  e2e %.1f ms (%d steps)" % (d["ms_per_step"], e.get("ms_per_step", float("nan")), e.get("steps", 0)))
PY
