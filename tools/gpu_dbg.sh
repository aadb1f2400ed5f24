#!/bin/bash
# in-graph cost of kernel families: step time with their work removed (timing experiments, wrong numerics)
for cfg in "0 0" "0 1" "7 0" "7 1"; do
  set -- $cfg
  UDA_B200_TC_DEBUG=$1 UDA_B200_BN_DEBUG=$2 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('TC_DEBUG=$1 BN_DEBUG=$2  ms/step', round(d['ms_per_step'],3))"
done
