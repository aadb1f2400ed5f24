#!/bin/bash
# quick confirmation after a kernel change: conv parity tests, conv micro-benchmark, training-step bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_conv_tc.log 2>&1; echo "== conv_tc + unet exit $? =="; grep -v "^E    +" gpurun_out/test_conv_tc.log | tail -n 5
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; tail -n 1 gpurun_out/conv_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], round(d['roofline']['achieved'],1), d['roofline']['traffic'], d['cpu_baseline']['value'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload adversarial --no-cpu-baseline > gpurun_out/bench_adv.log 2> gpurun_out/bench_adv.err; echo "== adversarial bench exit $? =="; tail -c 300 gpurun_out/bench_adv.log
