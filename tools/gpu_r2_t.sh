#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bench_shapes.py tests/test_gpu_stages.py tests/test_gpu_eval.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -12
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench_np.txt 2>&1; echo "== conv_bench exit $? =="; grep -E "dec2|dec3|dec4|head|layer1|weighted" gpurun_out/conv_bench_np.txt | cut -c1-200
for name in np nonp np2; do
  if [[ $name == nonp ]]; then export UDA_B200_HALO_NPACK=0; else unset UDA_B200_HALO_NPACK; fi
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['env'])
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
done
