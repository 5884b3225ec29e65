#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest staging"; timeout 900 python -m pytest -m gpu -q tests/test_gpu_host_staging.py --tb=short 2>&1 | grep -E "AssertionError|passed|failed|Error" | head
echo "== pytest ortho/shim"; timeout 900 python -m pytest -m gpu -q -x tests/test_shim.py tests/test_gpu_ortho.py tests/test_gpu_refsrc.py -k "not full_size" 2>&1 | tail -2
echo "== shim bench joint_10k"; timeout 900 python tools/shim_bench.py joint_10k 3 2>&1 | grep -v "driver staging" | tail -22
} > gpurun_out/r2_c26.log 2>&1
cut -c1-700 gpurun_out/r2_c26.log
