#!/bin/bash
# layer tests, adversarial bench, ncu launch list of one step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_layers.log 2>&1; echo "== layers exit $? =="; grep -v "^E    +" gpurun_out/test_layers.log | tail -n 5
timeout 600 python bench.py --steps 10 --warmup 3 --workload adversarial --no-cpu-baseline > gpurun_out/bench_adv.log 2> gpurun_out/bench_adv.err; echo "== adversarial bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_adv.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value']); print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 3 gpurun_out/bench_adv.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $? =="; wc -l gpurun_out/launches.csv
