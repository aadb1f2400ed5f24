#!/bin/bash
# round-2 diagnostics: ncu source-level captures of the dominant conv kernels, exported to CSV on the box
# (the .ncu-rep itself is too large to bring back)
mkdir -p gpurun_out
SHAPES="${SHAPES:-layer1 layer2 layer3 layer4 dec0.c1 l3.0 dec3.c1}"
timeout 300 python tools/ncu_conv.py $SHAPES > gpurun_out/ncu_conv_plain.log 2>&1; echo "== ncu_conv plain exit $? =="
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_" -o /tmp/r02_conv_full python tools/ncu_conv.py $SHAPES > gpurun_out/ncu_conv.log 2>&1
echo "== ncu exit $? =="; tail -n 2 gpurun_out/ncu_conv.log
ncu -i /tmp/r02_conv_full.ncu-rep --page raw --csv > gpurun_out/r02_conv_raw.csv 2> gpurun_out/raw.err; echo "raw export $?"
ncu -i /tmp/r02_conv_full.ncu-rep --page source --csv > gpurun_out/r02_conv_source.csv 2> gpurun_out/source.err; echo "source export $?"
ncu -i /tmp/r02_conv_full.ncu-rep --page details --csv > gpurun_out/r02_conv_details.csv 2> gpurun_out/details.err; echo "details export $?"
gzip -9 gpurun_out/r02_conv_source.csv gpurun_out/r02_conv_raw.csv gpurun_out/r02_conv_details.csv
ls -la gpurun_out
du -sm gpurun_out
