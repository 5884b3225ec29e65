#!/bin/bash
# N = 8: peer-push exchange on eight ranks (validation: sharded == undivided, checksum) + timing
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
{
nvidia-smi -L | wc -l
echo "== N=8 joint_10k peer push"; timeout 900 $TR --master-port 29711 bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | grep -E '^\{"metric|Error|error' | tail -3
} > gpurun_out/r2_c28.log 2>&1
python - <<'P'
import json
for l in open('gpurun_out/r2_c28.log'):
    if l.startswith('{"metric'):
        d = json.loads(l)
        print(d['n_gpus'], d['ms_per_step'], d['rank_ms_per_step'], d['roofline']['stage_ms'], d.get('sharded_equals_undivided'), d['checksum'], d.get('e2e',{}).get('ms_per_step'), d['config']['sharding'][-100:])
    else:
        print(l.rstrip()[:300])
P
