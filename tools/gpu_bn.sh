#!/bin/bash
mkdir -p gpurun_out
for f in layers conv_tc unet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -v "^E    +" gpurun_out/test_$f.log | tail -n 12
done
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline'] and round(d['roofline']['achieved'],1)); print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench.err
UDA_B200_FUSE_BN_BWD=0 timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_nofuse.log 2> gpurun_out/bench_nofuse.err; echo "== bench (no bn-bwd fusion) exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_nofuse.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')})
except Exception as e: print('bench parse failed', e)
PY
