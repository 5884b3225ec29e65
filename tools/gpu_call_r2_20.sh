#!/bin/bash
# two-level binning + flagged bucket order: hardware validation and A/B timing against the one-level binning
mkdir -p gpurun_out
{
echo "== pytest dsm/refsrc/pcl"; timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py tests/test_gpu_refsrc.py tests/test_ortho_from_pcl.py tests/test_gpu_smoke.py -k "not full_size" 2>&1 | tail -3
echo "== pytest dsm small windows"; AMB_DSM_PART_WINDOW=20000 timeout 900 python -m pytest -m gpu -q -x tests/test_gpu_dsm.py -k "not full_size" 2>&1 | tail -2
echo "== dsm stage timings, two-level"; timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -2
echo "== dsm stage timings, direct"; AMB_DSM_BINNING=direct timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1
for w in 8388608 50331648; do echo "== window $w"; AMB_DSM_PART_WINDOW=$w timeout 300 python tools/prof_run.py dsm 3 2>&1 | tail -1; done
echo "== launch list two-level"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:'dsm_|scan_' --csv --log-file gpurun_out/r2_c20_launches.csv python tools/prof_run.py dsm 1 > /dev/null 2>&1
python tools/ncu_launch_table.py gpurun_out/r2_c20_launches.csv
} > gpurun_out/r2_c20.log 2>&1
cat gpurun_out/r2_c20.log | cut -c1-400
