#!/bin/bash
# scheduling A/B: side-stream priority, programmatic dependent launch, weight-gradient stream, w_ft prefetch
mkdir -p gpurun_out
bench() { # name, env
  local name=$1; shift
  timeout 900 env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
}
bench base A=1
bench sideprio UDA_B200_WGRAD_PRIORITY=1
bench nopdl UDA_B200_PDL=0
bench sideprio_nopdl UDA_B200_WGRAD_PRIORITY=1 UDA_B200_PDL=0
bench nostream UDA_B200_WGRAD_STREAM=0
bench noprefetch UDA_B200_PREFETCH_WFT=0
bench base2 A=1
timeout 300 python -m pytest tests/test_gpu_unet.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -3
