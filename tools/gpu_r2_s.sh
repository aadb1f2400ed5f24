#!/bin/bash
# where does the halo kernel's epilogue time go on the 16-channel layers?  (experiment build, work-skipping switches)
for dbg in 0 1 8 9; do
  echo "== UDA_B200_TC_DEBUG=$dbg =="
  timeout 300 env UDA_B200_TC_DEBUG=$dbg python tools/trace_conv.py dec4.c2 head 2>&1 | grep -E "fwd|dgrad" | cut -c1-330
done
