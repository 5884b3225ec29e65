#!/bin/bash
# N = 2: peer-push halo exchange vs ncclSend/ncclRecv (hardware validation + timing)
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
{
nvidia-smi -L
echo "== pytest multi-rank"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_multi_rank.py 2>&1 | tail -15
echo "== N=2 joint_10k peer push"; timeout 600 $TR --master-port 29701 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep '^{"metric' | tail -1
echo "== N=2 joint_10k ncclSend/ncclRecv"; AMB_HALO_PEER=0 timeout 600 $TR --master-port 29702 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | grep '^{"metric' | tail -1
} > gpurun_out/r2_c22.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c22.log'):
    if l.startswith('{"metric'):
        d = json.loads(l)
        print(d['n_gpus'], d['ms_per_step'], d['rank_ms_per_step'], d['roofline']['stage_ms'], d.get('sharded_equals_undivided'), d['checksum'], d['config']['sharding'][-90:])
    else:
        print(l.rstrip()[:300])
P
