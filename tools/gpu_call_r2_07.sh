#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c07.log; : > $L
run() { echo "== $*" >> $L; timeout 900 "$@" 2>gpurun_out/r2_c07.err | tail -1 | cut -c1-6000 >> $L || echo "FAILED rc=$?" >> $L; tail -3 gpurun_out/r2_c07.err | cut -c1-400 >> $L; }
run python bench.py --workload joint_256 --steps 3 --warmup 1 --no-cpu-baseline
run python bench.py --workload dsm_256_holes --steps 3 --warmup 1 --no-cpu-baseline
run python bench.py --workload ortho_256_color --steps 3 --warmup 1 --no-cpu-baseline
run python bench.py --workload incremental_256 --steps 3 --warmup 1 --no-cpu-baseline
run python bench.py --steps 10 --warmup 3 --no-cpu-baseline
run python bench.py --workload dsm_c2 --steps 10 --warmup 3 --no-cpu-baseline
run python bench.py --workload dsm_c2_holes --steps 10 --warmup 3 --no-cpu-baseline
run python bench.py --workload ortho_c3_gray --steps 10 --warmup 3 --no-cpu-baseline
run python bench.py --workload ortho_c3_color --steps 10 --warmup 3 --no-cpu-baseline
run python bench.py --workload incremental_c5 --steps 5 --warmup 2 --no-cpu-baseline
python - <<'PY'
import json
for ln in open("gpurun_out/r2_c07.log"):
    if ln.startswith("=="): print(ln.strip()); continue
    try: d = json.loads(ln)
    except Exception: print("   ", ln.strip()[:300]); continue
    e = d.get("e2e") or {}
    print("    step %.3f ms  value %.3g  e2e %.1f ms  stages %s  checksum %s  extra %s" % (d["ms_per_step"], d["value"], e.get("ms_per_step", float("nan")), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, d.get("checksum"), {k: d[k] for k in ("incremental_equals_single_call", "dsm_cells_exact_path") if k in d}))
PY
