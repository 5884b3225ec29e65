#!/bin/bash
mkdir -p gpurun_out
{
echo "== pytest staging"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_host_staging.py 2>&1 | tail -40
echo "== pytest shim/ortho host"; timeout 900 python -m pytest -m gpu -q -x tests/test_shim.py tests/test_gpu_ortho.py tests/test_gpu_compact_mirrors.py tests/test_ortho_from_pcl.py 2>&1 | tail -4
echo "== e2e trace"; timeout 600 python tools/e2e_trace.py 2>&1 | tail -8
} > gpurun_out/r2_c24.log 2>&1
cut -c1-600 gpurun_out/r2_c24.log
