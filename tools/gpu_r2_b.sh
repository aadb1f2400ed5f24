#!/bin/bash
# round-2 iteration: micro-benchmark, role traces (experiment build), conv / stage parity tests, conv bench, step bench
mkdir -p gpurun_out
timeout 120 tools/exp/umma_rate 2>&1 | grep -E "issuers|---" > gpurun_out/umma_issuers.txt; cat gpurun_out/umma_issuers.txt
timeout 300 python tools/trace_conv.py layer1 layer2 layer3 layer4 dec0.c1 l3.0 dec3.c1 dec2.c1 > gpurun_out/trace_conv.txt 2>&1; echo "== trace exit $? =="; cut -c1-420 gpurun_out/trace_conv.txt
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bench_shapes.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_conv.log 2>&1; echo "== conv tests exit $? =="; tail -n 6 gpurun_out/test_conv.log
timeout 900 python -m pytest tests/test_gpu_stages.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/test_stages.log 2>&1; echo "== stage tests exit $? =="; grep -E "fwd|passed|failed|Error|assert" gpurun_out/test_stages.log | head -30
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], round(d['roofline']['achieved'],1), d['roofline']['frac'])
    print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 3 gpurun_out/bench.err
