#!/bin/bash
# fused conv + BatchNorm + activation: parity, then A/B in the bench and the in-graph timeline
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_bnfuse.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_bnfuse.log 2>&1; echo "== test_gpu_bnfuse exit $? =="; grep -v "^E    +" gpurun_out/test_bnfuse.log | tail -n 30
timeout 900 python -m pytest tests/test_gpu_unet.py tests/test_gpu_stages.py tests/test_gpu_adversary.py -q -m gpu -x --tb=short -p no:cacheprovider > gpurun_out/test_unet.log 2>&1; echo "== unet/stages/adversary exit $? =="; tail -n 5 gpurun_out/test_unet.log
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
bench fused A=1
bench unfused UDA_B200_FUSE_BN_APPLY=0
bench fused2 A=1
timeout 600 python tools/timeline.py --dump --out gpurun_out/r02_timeline_bnfuse.txt > gpurun_out/timeline.log 2>&1; echo "== timeline exit $? =="; head -8 gpurun_out/r02_timeline_bnfuse.txt
