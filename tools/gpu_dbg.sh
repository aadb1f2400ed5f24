#!/bin/bash
for a in 0 6 4; do
  echo "== UDA_B200_WGRAD_APG=$a =="
  UDA_B200_WGRAD_APG=$a timeout 300 python tools/conv_bench.py 2>&1 | awk -F'|' '{print $1, $4}' | grep -E "layer2|layer3|layer4|dec0|dec1|dec2.c1|l2.0|l3.0|l4.0|totals"
done
