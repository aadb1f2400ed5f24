#!/bin/bash
# losses parity + HBM micro-benchmark + the training-step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_losses.log 2>&1; echo "== losses exit $? =="; tail -n 12 gpurun_out/test_losses.log
timeout 600 python tools/hbm_bench.py > gpurun_out/hbm_stream.log 2>&1; echo "== hbm exit $? =="; grep -E "CE|consist|entropy|argmax" gpurun_out/hbm_stream.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline'] and round(d['roofline']['achieved'],1)); print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench.err
