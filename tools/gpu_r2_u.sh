#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu --tb=short -p no:cacheprovider -k "transposed or decoder_conv1 or stride2_4x4 or dec4.c1" 2>&1 | grep -v "^E    +" | tail -15
timeout 300 python tools/upconv_bench.py 2>&1 | tail -8
for name in up noup up2; do
  if [[ $name == noup ]]; then export UDA_B200_DOWNHALO=0; else unset UDA_B200_DOWNHALO; fi
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['kernel_breakdown_ms_per_step'].get('upconv_tc_fwd'), d['env'])
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
done
