#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest ortho"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_ortho_dominance.py tests/test_gpu_ortho.py tests/test_gpu_ortho_adversarial.py tests/test_gpu_refsrc.py tests/test_gpu_compact_mirrors.py tests/test_shim.py tests/test_gpu_host_staging.py 2>&1 | tail -3
for wl in joint_10k ortho_c3_gray ortho_c3_color incremental_c5; do
  echo "== bench $wl"; timeout 900 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
done
} > gpurun_out/r2_c33.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c33.log'):
    if l.startswith('{"metric'):
        d = json.loads(l); print(d['config']['workload'], round(d['ms_per_step'],3), {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, 'e2e', round(d['e2e']['ms_per_step'],2), d.get('incremental_equals_single_call'), d['checksum'])
    else:
        print(l.rstrip()[:300])
P
