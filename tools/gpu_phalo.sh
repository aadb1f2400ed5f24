#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider -k "pitched" > gpurun_out/test_phalo.log 2>&1; echo "== phalo tests exit $? =="; grep -v "^E    +" gpurun_out/test_phalo.log | tail -n 30
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log | awk -F'|' '{print $1, $2, $3}'
