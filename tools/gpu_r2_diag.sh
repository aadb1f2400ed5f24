#!/bin/bash
# round-2 diagnostics: tensor-pipe / TMA micro-benchmarks, then ncu source-level captures of the dominant conv kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 120 tools/exp/umma_rate > gpurun_out/umma_rate.txt 2>&1; echo "== umma_rate exit $? =="; cat gpurun_out/umma_rate.txt
timeout 300 python tools/ncu_conv.py > gpurun_out/ncu_conv_plain.log 2>&1; echo "== ncu_conv plain exit $? =="; tail -n 3 gpurun_out/ncu_conv_plain.log
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_" -o gpurun_out/r02_conv_full python tools/ncu_conv.py > gpurun_out/ncu_conv.log 2>&1
echo "== ncu exit $? =="; tail -n 3 gpurun_out/ncu_conv.log; ls -la gpurun_out/*.ncu-rep
