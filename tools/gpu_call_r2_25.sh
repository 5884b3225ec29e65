#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest staging"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_host_staging.py 2>&1 | tail -3
echo "== shim bench joint_10k"; timeout 900 python tools/shim_bench.py joint_10k 3 2>&1 | tail -40
echo "== shim bench joint_10k, 32 staging threads"; AMB_STAGING_THREADS=32 timeout 900 python tools/shim_bench.py joint_10k 3 2>&1 | grep -v "driver staging" | tail -24
} > gpurun_out/r2_c25.log 2>&1
cut -c1-700 gpurun_out/r2_c25.log
