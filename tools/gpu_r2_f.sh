#!/bin/bash
mkdir -p gpurun_out
for f in bench_shapes stages unet layers conv_tc; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider -s > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -E "passed|failed|^E  |eval logits|fwd .*L2" gpurun_out/test_$f.log | head -24
done
bench() { # name, extra args / env
  local name=$1; shift
  timeout 900 env "$@" python bench.py --steps 10 --warmup 3 --no-cpu-baseline $BARGS > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))
    for k in ('adversarial','adversarial_grl'):
        if k in d: print(' ', k, {q:round(d[k][q],2) for q in ('value','ms_per_step')}, 'e2e', round(d[k]['e2e']['value'],1))
    print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
}
BARGS="--no-sub" bench fused A=1
BARGS="--no-sub" bench unfused UDA_B200_FUSE_UPCAT=0
timeout 600 python tools/eval_bench.py > gpurun_out/eval_bench.log 2>&1; echo "== eval_bench exit $? =="; tail -2 gpurun_out/eval_bench.log | cut -c1-400
