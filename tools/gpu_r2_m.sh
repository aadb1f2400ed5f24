#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bnfuse.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -6
timeout 600 python tools/trace_step.py > gpurun_out/r02_trace_step.txt 2> gpurun_out/trace_step.err; echo "== trace_step fused exit $? =="; tail -3 gpurun_out/trace_step.err; grep -A1 "persist+bn" gpurun_out/r02_trace_step.txt | head -30 | cut -c1-200
for name in a b; do
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
done
