#!/bin/bash
mkdir -p gpurun_out
for f in layers conv_tc unet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -v "^E    +" gpurun_out/test_$f.log | tail -n 6
done
for pdl in 1 0; do
  UDA_B200_PDL=$pdl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>gpurun_out/bench_pdl$pdl.err | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PDL=$pdl  ms/step', round(d['ms_per_step'],3), 'img/s', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))
except Exception as e: print('PDL=$pdl failed', e)"
  tail -n 3 gpurun_out/bench_pdl$pdl.err
done
