#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest dsm (f32 default)"; timeout 600 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py -k "not full_size" 2>&1 | tail -2
echo "== dsm stage timings f32"; timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1
echo "== dsm stage timings f32 bps4"; AMB_DSM_F32_BPS=4 timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1
} > gpurun_out/r2_c19.log 2>&1
cat gpurun_out/r2_c19.log | cut -c1-330
