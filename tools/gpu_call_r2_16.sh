#!/bin/bash
# final-state evidence: ncu --set full of the two dominant kernels at joint_10k, launch list of the bench command
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:dsm_gather_kernel_f32 -c 1 -o gpurun_out/r2_final_gather_f32_joint -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_c16_ncu1.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ortho_kernel_dom -c 1 -o gpurun_out/r2_final_ortho_dom_joint -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_c16_ncu2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2_c16_ncu3.log 2>&1
timeout 600 python bench.py --steps 10 --warmup 3 2>gpurun_out/r2_c16.err | tail -1 > gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/r2_c16.err | tail -1 > gpurun_out/r2_bench_reference_arm.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_n1.json")); r = json.load(open("gpurun_out/r2_bench_reference_arm.json"))
print("ours  step %.3f ms value %.3g e2e %.1f ms" % (d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"]))
print("ref   %.3g cells/s on %d cores (%s)" % (r["value"], r["cpu_baseline"]["cores"], r["cpu_baseline"]["kind"]))
print("e2e ratio %.0f  device ratio %.0f" % (d["e2e"]["value"] / r["value"], d["value"] / r["value"]))
PY
