#!/bin/bash
# round 2, call 2: f32 gather on hardware — parity in both precisions, stage timings, ncu captures
mkdir -p gpurun_out
{
echo "== pytest f32 (default)"; timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py tests/test_gpu_smoke.py tests/test_shim.py tests/test_ortho_from_pcl.py 2>&1 | tail -4
echo "== pytest f64"; AMB_DSM_PRECISION=f64 timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py tests/test_shim.py 2>&1 | tail -4
echo "== dsm stage timings f32"; timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -3
echo "== dsm stage timings f64"; AMB_DSM_PRECISION=f64 timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -3
echo "== bench default"; timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1
} > gpurun_out/r2_c02.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_gather_kernel_f32 -c 1 -o gpurun_out/r2_gather_f32 -f python tools/prof_run.py dsm 1 > gpurun_out/r2_c02_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ortho_kernel_dom -c 1 -o gpurun_out/r2_ortho_dom_joint -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_c02_ncu2.log 2>&1
cat gpurun_out/r2_c02.log
