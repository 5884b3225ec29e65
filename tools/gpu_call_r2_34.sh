#!/bin/bash
# orthomosaic split in two kernels (cull lists, selection): validation + A/B against the fused kernel
mkdir -p gpurun_out
{
echo "== pytest ortho, split kernels"; AMB_ORTHO_SPLIT=1 timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_ortho_dominance.py tests/test_gpu_ortho.py tests/test_gpu_ortho_adversarial.py tests/test_gpu_refsrc.py tests/test_gpu_compact_mirrors.py 2>&1 | tail -2
for sp in 0 1; do for wl in joint_10k ortho_c3_gray incremental_c5; do
  echo "== split=$sp bench $wl"; AMB_ORTHO_SPLIT=$sp timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | tail -1
done; done
} > gpurun_out/r2_c34.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c34.log'):
    if l.startswith('{"metric'):
        d = json.loads(l); print(d['config']['workload'], round(d['ms_per_step'],3), 'ortho', round(d['roofline']['stage_ms']['ortho'],3), d.get('incremental_equals_single_call'), d['checksum'])
    else:
        print(l.rstrip()[:300])
P
