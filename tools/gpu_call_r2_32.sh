#!/bin/bash
# final captures of the round: ncu --set full of the four dominant kernels at joint_10k, launch list of a bench run, bench line
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'dsm_partition_kernel|dsm_fine_scatter_kernel|dsm_gather_kernel_f32|ortho_kernel_dom' -c 4 -f -o gpurun_out/r2_final2_joint python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_final2_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 900 python bench.py > gpurun_out/r2_final2_bench_n1.json 2> gpurun_out/r2_final2_bench_n1.err
tail -c 3000 gpurun_out/r2_final2_bench_n1.json
ls -la gpurun_out | tail -8
