#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/trace_conv.py layer2 layer3 dec2.c1 > gpurun_out/trace_conv.txt 2>&1; echo "== trace exit $? =="; grep wgrad gpurun_out/trace_conv.txt | cut -c1-420
for f in conv_tc bench_shapes unet layers; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -E "passed|failed|^E  |eval logits|bf16 discriminator" gpurun_out/test_$f.log | head -20
done
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log
timeout 600 python tools/eval_bench.py > gpurun_out/eval_bench.log 2>&1; echo "== eval_bench exit $? =="; tail -5 gpurun_out/eval_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], round(d['roofline']['achieved'],1), d['roofline']['frac'])
    print('adversarial', {k:d['adversarial'][k] for k in ('value','ms_per_step')}, d['adversarial']['e2e']['value'])
    print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench.err
