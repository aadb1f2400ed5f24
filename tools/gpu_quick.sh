#!/bin/bash
# quick A/B of the tensor-core convolution paths
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_conv_tc.log 2>&1; echo "== conv_tc (persistent) exit $? =="; tail -n 3 gpurun_out/test_conv_tc.log
UDA_B200_TC_PERSIST=0 timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_conv_tc_np.log 2>&1; echo "== conv_tc (non-persistent) exit $? =="; tail -n 3 gpurun_out/test_conv_tc_np.log
UDA_B200_TC_HALO=0 timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_conv_tc_nohalo.log 2>&1; echo "== conv_tc (no halo) exit $? =="; tail -n 2 gpurun_out/test_conv_tc_nohalo.log
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_unet.log 2>&1; echo "== unet exit $? =="; tail -n 3 gpurun_out/test_unet.log
timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['value'], d['roofline'] and round(d['roofline']['achieved'],1)); print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
tail -n 5 gpurun_out/bench.err
