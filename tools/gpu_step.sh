#!/bin/bash
# conv tile-shape tests + conv micro-benchmark + the training-step bench + ncu launch list of one step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_conv_tc.log 2>&1; echo "== conv_tc exit $? =="; grep -v "^E    +" gpurun_out/test_conv_tc.log | tail -n 5
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log | awk -F'|' '{print $1, $2, $3}'
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline'] and round(d['roofline']['achieved'],1)); print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench.err
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1500 -c 330 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $? =="; tail -n 2 gpurun_out/ncu_launch.log; wc -l gpurun_out/launches.csv
