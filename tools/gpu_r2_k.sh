#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bnfuse.py -q -m gpu --tb=short -p no:cacheprovider 2>&1 | tail -3
timeout 600 env UDA_B200_FUSE_BN_APPLY=1 python tools/trace_step.py > gpurun_out/r02_trace_step_bnfuse.txt 2> gpurun_out/trace_step.err; echo "== trace_step fused exit $? =="; tail -5 gpurun_out/trace_step.err; grep -A1 "+bn" gpurun_out/r02_trace_step_bnfuse.txt | head -60; tail -9 gpurun_out/r02_trace_step_bnfuse.txt
