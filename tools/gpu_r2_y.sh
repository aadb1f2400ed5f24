#!/bin/bash
# last run of the round: the driver's own GPU test command, smoke, the default bench line, CUPTI timeline, ncu launch list
R=${R:-r02}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/test_all.log 2>&1; echo "== pytest -m gpu exit $? =="; tail -n 3 gpurun_out/test_all.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -n 1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -n 2 gpurun_out/bench.err
timeout 200 python tools/timeline.py --dump --out gpurun_out/${R}_timeline.txt > gpurun_out/timeline.log 2>&1; echo "== timeline exit $? =="; head -n 4 gpurun_out/${R}_timeline.txt
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub"
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv --log-file /tmp/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $? =="; wc -l /tmp/launches.csv
python tools/launch_shares.py /tmp/launches.csv gpurun_out/${R}_launch_shares.csv --step > /dev/null; head -n 8 gpurun_out/${R}_launch_shares.csv
python - gpurun_out/${R}_bench.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print({k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), "roofline", {k:(d.get("roofline") or {}).get(k) for k in ("achieved","frac","ms_per_step")})
print({k:(v.get('value'),v.get('ms_per_step'),(v.get('e2e') or {}).get('value')) for k,v in d.items() if isinstance(v,dict) and 'ms_per_step' in v and k!='roofline'})
PY
