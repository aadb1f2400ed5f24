#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_layers.py tests/test_gpu_eval.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_layers.log 2>&1; echo "== layers+eval exit $? =="; grep -v "^E    +" gpurun_out/test_layers.log | tail -n 6
timeout 600 python tools/hbm_bench.py 2>&1 | grep -E "maxpool|upcat" 
timeout 600 python tools/eval_bench.py > gpurun_out/eval_bench.log 2>&1; echo "== eval_bench exit $? =="; tail -n 3 gpurun_out/eval_bench.log
