#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_losses.py tests/test_gpu_adversary.py tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | grep -v "^E    +" | tail -8
for v in tma rows; do
  echo "== hbm_bench loss kernels, $v =="
  if [[ $v == tma ]]; then unset UDA_B200_LOSS_TMA; else export UDA_B200_LOSS_TMA=0; fi
  timeout 300 python tools/hbm_bench.py 2>&1 | grep -E "CE|consistency|entropy"
done
