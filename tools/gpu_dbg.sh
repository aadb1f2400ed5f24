#!/bin/bash
mkdir -p gpurun_out
for d in 0 7 2 4 1; do
  UDA_B200_TC_DEBUG=$d UDA_B200_TC_PHALO=0 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); b=d['kernel_breakdown_ms_per_step']
print('debug=$d  ms/step', round(d['ms_per_step'],3), ' fwd', b.get('conv2d_tc_fwd'), ' dgrad', b.get('conv2d_tc_dgrad'), ' wgrad', b.get('conv2d_tc_wgrad'))"
done
