#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "pitched" > gpurun_out/test_phalo.log 2>&1; echo "== phalo tests exit $? =="; grep -v "^E    +" gpurun_out/test_phalo.log | tail -n 10
UDA_B200_TC_PHALO=2 timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench_phalo2.log 2>&1; echo "== conv_bench (phalo forced) exit $? =="; cat gpurun_out/conv_bench_phalo2.log | awk -F'|' '{print $1, $2, $3}' | grep -E "layer2|layer3|dec0|dec1|totals"
