#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bnfuse.py tests/test_gpu_layers.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -15
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_stages.py -q -m gpu -x --tb=short -p no:cacheprovider 2>&1 | tail -4
bench() { # name, env
  local name=$1; shift
  timeout 900 env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))
    print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
}
bench merged UDA_B200_FUSE_BN_APPLY=1
bench twolaunch UDA_B200_FUSE_BN_APPLY=1 UDA_B200_BN_BWD_MERGED=0
bench merged2 UDA_B200_FUSE_BN_APPLY=1
