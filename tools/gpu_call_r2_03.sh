#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest f32 (default)"; timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py tests/test_gpu_smoke.py tests/test_shim.py 2>&1 | tail -4
echo "== pytest f64"; AMB_DSM_PRECISION=f64 timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py tests/test_shim.py 2>&1 | tail -4
echo "== dsm stage timings f32"; timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -2
echo "== dsm stage timings f64"; AMB_DSM_PRECISION=f64 timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -2
} > gpurun_out/r2_c03.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_gather_kernel_f32 -c 1 -o gpurun_out/r2_gather_f32_v2 -f python tools/prof_run.py dsm 1 > gpurun_out/r2_c03_ncu1.log 2>&1
cat gpurun_out/r2_c03.log
