#!/bin/bash
# final state of the round: whole GPU suite, smoke, default bench line
mkdir -p gpurun_out
{
echo "== pytest -m gpu (all)"; timeout 1500 python -m pytest -m gpu -q -x tests 2>&1 | tail -3
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
} > gpurun_out/r2_c35.log 2>&1
timeout 900 python bench.py > gpurun_out/r2_final3_bench_n1.json 2> gpurun_out/r2_final3_bench_n1.err
cat gpurun_out/r2_c35.log
python - <<'P'
import json
d = json.loads(open('gpurun_out/r2_final3_bench_n1.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['e2e']['ms_per_step'], d['checksum'], [ (k['kernel'][:30], round(k['frac'],3)) for k in d['roofline']['kernels']], d['cpu_baseline']['value'])
P
