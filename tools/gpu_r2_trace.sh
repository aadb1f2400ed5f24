#!/bin/bash
# round-2: per-CTA role traces of the conv kernels (experiment build) + the new bench-shape parity tests + CPU-side fixes on GPU
mkdir -p gpurun_out
timeout 300 python tools/trace_conv.py layer1 layer2 layer3 layer4 dec0.c1 l3.0 dec3.c1 dec2.c1 > gpurun_out/trace_conv.txt 2>&1; echo "== trace exit $? =="; cat gpurun_out/trace_conv.txt
timeout 900 python -m pytest tests/test_gpu_bench_shapes.py -q -m gpu --tb=short -p no:cacheprovider -x > gpurun_out/test_bench_shapes.log 2>&1; echo "== bench shapes exit $? =="; tail -n 15 gpurun_out/test_bench_shapes.log
