#!/bin/bash
# Runs every GPU test file in its own process (a CUDA fault in one file cannot poison the others);
# logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
for f in losses eval layers conv_tc unet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; tail -n 4 gpurun_out/test_$f.log
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -n 3 gpurun_out/smoke.log
if [ "$1" == "bench" ]; then
  timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
fi
