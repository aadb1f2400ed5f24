#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_conv.py layer2 layer3 layer4 dec0.c1 dec2.c1 > gpurun_out/trace_conv.txt 2>&1; echo "== trace exit $? =="; grep wgrad gpurun_out/trace_conv.txt | cut -c1-420
UDA_B200_TC_PHALO=2 timeout 300 python tools/trace_conv.py layer2 dec1.c1 > gpurun_out/trace_conv_phalo64.txt 2>&1; echo "== trace phalo64 exit $? =="; grep -v wgrad gpurun_out/trace_conv_phalo64.txt | cut -c1-420
timeout 900 python -m pytest tests/test_gpu_stages.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/test_stages.log 2>&1; echo "== stage tests exit $? =="; grep -E "fwd|passed|failed|Error|assert" gpurun_out/test_stages.log | head -30
UDA_B200_TC_PHALO=2 timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench_phalo64.log 2>&1; echo "== conv_bench phalo64 exit $? =="; grep -E "layer2|dec1" gpurun_out/conv_bench_phalo64.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], round(d['roofline']['achieved'],1), d['roofline']['frac'])
    print('adversarial', {k:d['adversarial'][k] for k in ('value','ms_per_step')}, d['adversarial']['e2e']['value'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --workload finetune > gpurun_out/bench_ft.log 2> gpurun_out/bench_ft.err; echo "== bench finetune exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_ft.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], round(d['roofline']['achieved'],1), d['roofline']['frac'])
    print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench_ft.err
