#!/bin/bash
# trig-free dominance prologue (ortho), binning windows / in-flight depth, e2e trace
mkdir -p gpurun_out
{
echo "== pytest ortho"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_ortho_dominance.py tests/test_gpu_ortho.py tests/test_gpu_ortho_adversarial.py tests/test_gpu_refsrc.py tests/test_gpu_compact_mirrors.py -k "not full_size" 2>&1 | tail -3
echo "== bench joint_10k"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
for w in 50331648 67108864 100663296; do echo "== window $w"; AMB_DSM_PART_WINDOW=$w timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1; done
for f in 1 4; do echo "== inflight $f (window 48M)"; AMB_DSM_PART_WINDOW=50331648 AMB_DSM_FINE_INFLIGHT=$f timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1; done
echo "== e2e trace"; timeout 600 python tools/e2e_trace.py 2>&1 | tail -8
} > gpurun_out/r2_c21.log 2>&1
cat gpurun_out/r2_c21.log | cut -c1-1500
