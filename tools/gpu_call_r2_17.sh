#!/bin/bash
# ortho kernel tuning variants (cells per thread x resident blocks per SM): correctness + timing
mkdir -p gpurun_out
L=gpurun_out/r2_c18.log; : > $L
for v in default s4b3i32 s4b2i64 s4b1i128; do
  if [ $v = default ]; then unset AMB_LIB_PATH; else export AMB_LIB_PATH=$PWD/build/variants/libamb_$v.so; fi
  echo "== $v" >> $L
  timeout 300 python -m pytest -m gpu -q -x tests/test_gpu_ortho.py tests/test_gpu_ortho_adversarial.py -k "not large" 2>&1 | tail -1 >> $L
  for w in ortho_c3_gray joint_10k; do
    timeout 300 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print('   %-14s step %.3f ms ortho %.3f ms' % (d['config']['workload'], d['ms_per_step'], d['roofline']['stage_ms']['ortho']))" >> $L
  done
done
cat $L
