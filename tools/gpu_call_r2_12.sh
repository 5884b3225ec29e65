#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r2_c12.log; : > $L
echo "== pytest ortho (adversarial, parity, dominance, refsrc incl. full-size C3 stripe)" >> $L
timeout 900 python -m pytest -m gpu -q tests/test_gpu_ortho_adversarial.py tests/test_gpu_ortho.py tests/test_gpu_ortho_dominance.py tests/test_gpu_refsrc.py tests/test_shim.py 2>&1 | tail -8 >> $L
echo "== bench joint_10k" >> $L
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('step %.3f ms e2e %.1f ms stages %s' % (d['ms_per_step'], d['e2e']['ms_per_step'], {k: round(v, 3) for k, v in d['roofline']['stage_ms'].items()}))" >> $L
echo "== bench ortho_c3_gray" >> $L
timeout 600 python bench.py --workload ortho_c3_gray --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('step %.3f ms stages %s' % (d['ms_per_step'], {k: round(v, 3) for k, v in d['roofline']['stage_ms'].items()}))" >> $L
cat $L
