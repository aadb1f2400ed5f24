#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_conv_tc.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_conv_tc.log 2>&1; echo "== conv_tc exit $? =="; tail -n 15 gpurun_out/test_conv_tc.log
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log
UDA_B200_WGRAD_BIG=0 timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench_nobig.log 2>&1; echo "== conv_bench (no big) exit $? =="; tail -n 1 gpurun_out/conv_bench_nobig.log
timeout 300 python tools/loss_one.py ce > gpurun_out/plain.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"seg_loss" -s 2 -c 1 -o gpurun_out/prof_loss python tools/loss_one.py ce > gpurun_out/ncu_loss.log 2>&1
echo "== ncu exit $? =="; tail -n 3 gpurun_out/ncu_loss.log
