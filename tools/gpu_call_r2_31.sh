#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
echo "== pytest multi-rank"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_multi_rank.py 2>&1 | tail -3
echo "== N=2 joint_10k peer push"; timeout 600 $TR --master-port 29741 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep '^{"metric' | tail -1
} > gpurun_out/r2_c31.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c31.log'):
    if l.startswith('{"metric'):
        d = json.loads(l)
        print(d['config']['workload'], d['n_gpus'], round(d['ms_per_step'],3), [round(x,3) for x in d['rank_ms_per_step']], 'inflight', d['stage_ms_last_timed_step'], d.get('sharded_equals_undivided'), d['checksum'])
    else:
        print(l.rstrip()[:300])
P
