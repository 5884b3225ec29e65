#!/bin/bash
mkdir -p gpurun_out
N=${1:-2}
L=gpurun_out/r2_c09_n$N.log; : > $L
run() { echo "== $*" >> $L; timeout 900 "$@" 2>gpurun_out/r2_c09.err | tail -1 | cut -c1-6000 >> $L || echo "FAILED rc=$?" >> $L; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$" gpurun_out/r2_c09.err | tail -4 | cut -c1-500 >> $L; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
run $TR bench.py --gpus $N --workload joint_1k --steps 3 --warmup 1 --no-cpu-baseline --no-e2e
run env AMB_HALO_EXCHANGE=1 $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-verify
run $TR bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline
python - $L <<'PY'
import json, sys
for ln in open(sys.argv[1]):
    if ln.startswith("=="): print(ln.strip()[:200]); continue
    try: d = json.loads(ln)
    except Exception: print("   ", ln.strip()[:400]); continue
    e = d.get("e2e") or {}
    print("    step %.3f ms  value %.3g  e2e %.1f ms  stages %s  checksum %s  ranks %s extra %s" % (d["ms_per_step"], d["value"], e.get("ms_per_step", float("nan")), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, d.get("checksum"), [round(x,3) for x in d.get("rank_ms_per_step",[])], {k: d[k] for k in ("incremental_equals_single_call", "sharded_equals_undivided") if k in d}))
PY
