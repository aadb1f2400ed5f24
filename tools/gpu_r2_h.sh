#!/bin/bash
# session 2 of round 2, first call: in-graph kernel timeline (CUPTI), A/B of the high-priority capture stream
mkdir -p gpurun_out
timeout 600 python tools/timeline.py --dump --out gpurun_out/r02_timeline.txt > gpurun_out/timeline.log 2>&1; echo "== timeline exit $? =="; head -60 gpurun_out/r02_timeline.txt
bench() { # name, extra args / env
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
bench prio A=1
bench noprio UDA_B200_CAPTURE_PRIORITY=0
bench prio2 A=1
bench noprio2 UDA_B200_CAPTURE_PRIORITY=0
bench bnbwd UDA_B200_FUSE_BN_BWD=1
