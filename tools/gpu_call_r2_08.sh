#!/bin/bash
# 2 GPUs: the in-library exchange step (NCCL loaded by the library), sharded == undivided, scaling
mkdir -p gpurun_out
L=gpurun_out/r2_c08.log; : > $L
run() { echo "== $*" >> $L; timeout 900 "$@" 2>gpurun_out/r2_c08.err | tail -1 | cut -c1-6000 >> $L || echo "FAILED rc=$?" >> $L; tail -4 gpurun_out/r2_c08.err | cut -c1-500 >> $L; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
run $TR bench.py --gpus 2 --workload joint_1k --steps 3 --warmup 1 --no-cpu-baseline
run $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline
run $TR bench.py --gpus 2 --workload ortho_c3_gray --steps 10 --warmup 3 --no-cpu-baseline --no-e2e
run python bench.py --workload incremental_256 --steps 3 --warmup 1 --no-cpu-baseline --no-e2e
python - <<'PY'
import json
for ln in open("gpurun_out/r2_c08.log"):
    if ln.startswith("=="): print(ln.strip()); continue
    try: d = json.loads(ln)
    except Exception: print("   ", ln.strip()[:400]); continue
    e = d.get("e2e") or {}
    print("    step %.3f ms  value %.3g  e2e %.1f ms  stages %s  checksum %s  ranks %s extra %s" % (d["ms_per_step"], d["value"], e.get("ms_per_step", float("nan")), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, d.get("checksum"), [round(x,3) for x in d.get("rank_ms_per_step",[])], {k: d[k] for k in ("incremental_equals_single_call", "sharded_equals_undivided") if k in d}))
PY
