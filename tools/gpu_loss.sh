#!/bin/bash
# loss kernels: parity tests with the streaming path on and off, then the HBM micro-benchmark of the variants
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_losses.log 2>&1; echo "== losses (stream) exit $? =="; tail -n 12 gpurun_out/test_losses.log
UDA_B200_LOSS_PPT=2 timeout 600 python -m pytest tests/test_gpu_losses.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_losses_p2.log 2>&1; echo "== losses (stream ppt2) exit $? =="; tail -n 3 gpurun_out/test_losses_p2.log
UDA_B200_LOSS_PPT=1 timeout 600 python -m pytest tests/test_gpu_losses.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_losses_p1.log 2>&1; echo "== losses (stream ppt1) exit $? =="; tail -n 3 gpurun_out/test_losses_p1.log
UDA_B200_LOSS_PPT=1 timeout 600 python tools/hbm_bench.py > gpurun_out/hbm_stream_p1.log 2>&1; echo "== hbm (stream ppt1) exit $? =="; grep -E "CE|consist|entropy" gpurun_out/hbm_stream_p1.log
UDA_B200_LOSS_PPT=2 timeout 600 python tools/hbm_bench.py > gpurun_out/hbm_stream_p2.log 2>&1; echo "== hbm (stream ppt2) exit $? =="; grep -E "CE|consist|entropy" gpurun_out/hbm_stream_p2.log
