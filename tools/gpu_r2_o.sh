#!/bin/bash
# confirmation of the defaults: the whole GPU suite (as the driver runs it) + smoke, then the default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "== bench default exit $? =="
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_default.log').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['roofline']['traffic_note'][100:170], 'cpu', d['cpu_baseline']['value'])
PY
tail -n 3 gpurun_out/bench_default.err
