#!/bin/bash
# ncu capture of the HBM-bound kernels inside one bench run (after the same command exited 0 without ncu)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --set full --clock-control none --import-source on \
  -k regex:"bn_apply_kernel|bn_stats_kernel|bn_bwd_reduce_kernel|bn_bwd_apply_kernel|seg_loss_kernel|upcat_fwd_kernel|upcat_bwd_kernel|maxpool_bwd_kernel|adam_kernel|nchw_to_nhwc_smallc" \
  -s 600 -c 260 -o gpurun_out/prof_hbm $CMD > gpurun_out/ncu_hbm.log 2>&1
echo "== ncu hbm exit $? =="; tail -n 2 gpurun_out/ncu_hbm.log
