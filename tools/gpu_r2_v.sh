#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/timeline.py --dump --out gpurun_out/r02_timeline.txt > gpurun_out/timeline.log 2>&1; echo "== timeline exit $? =="; head -30 gpurun_out/r02_timeline.txt
timeout 600 python tools/conv_bench.py > gpurun_out/r02_conv_bench.txt 2>&1; echo "== conv_bench exit $? =="; tail -22 gpurun_out/r02_conv_bench.txt | cut -c1-200
