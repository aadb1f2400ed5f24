#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/trace_step.py > gpurun_out/r02_trace_step.txt 2> gpurun_out/trace_step.err; echo "== trace_step exit $? =="; tail -5 gpurun_out/trace_step.err; head -5 gpurun_out/r02_trace_step.txt; tail -8 gpurun_out/r02_trace_step.txt
