#!/bin/bash
# confirmation of the defaults: all GPU tests + smoke, bench (default switches), 4-warp-epilogue build is not shipped
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider -x 2>&1 | grep -v "^E    +" | tail -8
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for name in default default2 nobnfuse; do
  if [[ $name == nobnfuse ]]; then export UDA_B200_FUSE_BN_APPLY=0; else unset UDA_B200_FUSE_BN_APPLY; fi
  timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['env'])
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
done
