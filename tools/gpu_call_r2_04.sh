#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest f32 (default)"; timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py tests/test_gpu_smoke.py 2>&1 | tail -3
echo "== dsm stage timings f32 bps5"; timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1
echo "== dsm stage timings f32 bps4"; AMB_DSM_F32_BPS=4 timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1
} > gpurun_out/r2_c06.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_gather_kernel_f32 -c 1 -o gpurun_out/r2_gather_f32_v5 -f python tools/prof_run.py dsm 1 > gpurun_out/r2_c06_ncu1.log 2>&1
cat gpurun_out/r2_c06.log
