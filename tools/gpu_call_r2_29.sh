#!/bin/bash
# N = 2 diagnosis: stage times of the last step of the back-to-back region, host enqueue time per step
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
echo "== N=2 joint_10k peer push"; timeout 600 $TR --master-port 29721 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep '^{"metric' | tail -1
echo "== N=2 joint_10k ncclSend/ncclRecv"; AMB_HALO_PEER=0 timeout 600 $TR --master-port 29722 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --no-verify 2>&1 | grep '^{"metric' | tail -1
echo "== N=1 joint_10k"; timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep '^{"metric' | tail -1
} > gpurun_out/r2_c29.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c29.log'):
    if l.startswith('{"metric'):
        d = json.loads(l)
        print(d['n_gpus'], round(d['ms_per_step'],3), [round(x,3) for x in d['rank_ms_per_step']], 'sync', {k: round(v,3) for k,v in d['roofline']['stage_ms'].items()}, 'inflight', {k: round(v,3) for k,v in d['stage_ms_last_timed_step'].items()}, 'enq', round(d['host_enqueue_ms_per_step'],3), d.get('sharded_equals_undivided'))
    else:
        print(l.rstrip()[:300])
P
