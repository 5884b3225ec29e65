#!/bin/bash
# pageable staging + shared shim backend: validation, shim path timed at the benchmark size
mkdir -p gpurun_out
{
echo "== pytest staging/shim/ortho host"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_host_staging.py tests/test_shim.py tests/test_gpu_ortho.py tests/test_gpu_compact_mirrors.py tests/test_ortho_from_pcl.py 2>&1 | tail -4
echo "== shim bench joint_10k"; timeout 900 python tools/shim_bench.py joint_10k 4 2>&1 | tail -4
echo "== bench joint_10k"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/r2_c23.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c23.log'):
    if l.startswith('{"metric'):
        d = json.loads(l); print(d['ms_per_step'], d['roofline']['stage_ms'], d['e2e'])
    else:
        print(l.rstrip()[:900])
P
