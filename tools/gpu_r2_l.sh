#!/bin/bash
# 8-warp epilogue: parity, bench A/B (fused BatchNorm apply on/off), in-graph traces
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_conv_tc.py tests/test_gpu_bench_shapes.py tests/test_gpu_bnfuse.py tests/test_gpu_layers.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_conv.log 2>&1; echo "== conv tests exit $? =="; grep -v "^E    +" gpurun_out/test_conv.log | tail -n 12
bench() { # name, env
  local name=$1; shift
  timeout 900 env "$@" python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/bench_$name.log 2> gpurun_out/bench_$name.err; echo "== bench $name exit $? =="
  python - "$name" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step')}, 'e2e', round(d['e2e']['value'],1), 'roof', round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))
    print(d['kernel_breakdown_ms_per_step'])
except Exception as e: print('bench parse failed', e)
PY
  tail -n 3 gpurun_out/bench_$name.err
}
bench epi8 A=1
bench epi8_bnfuse UDA_B200_FUSE_BN_APPLY=1
timeout 600 python tools/trace_step.py > gpurun_out/r02_trace_step_epi8.txt 2> gpurun_out/trace_step.err; echo "== trace_step exit $? =="; tail -8 gpurun_out/r02_trace_step_epi8.txt
timeout 600 env UDA_B200_FUSE_BN_APPLY=1 python tools/trace_step.py > gpurun_out/r02_trace_step_epi8_bnfuse.txt 2> gpurun_out/trace_step.err; echo "== trace_step fused exit $? =="; grep -A1 "phalo+bn" gpurun_out/r02_trace_step_epi8_bnfuse.txt | head -8; tail -9 gpurun_out/r02_trace_step_epi8_bnfuse.txt
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench_epi8.txt 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench_epi8.txt | tail -22
