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
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -c 3000 gpurun_out/bench.log; tail -n 5 gpurun_out/bench.err
  timeout 600 python bench.py --steps 10 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_eager.log 2> gpurun_out/bench_eager.err; echo "== eager bench exit $? =="; tail -c 600 gpurun_out/bench_eager.log | cut -c1-400; tail -n 5 gpurun_out/bench_eager.err
  timeout 600 python bench.py --steps 10 --warmup 3 --workload adversarial --no-cpu-baseline > gpurun_out/bench_adv.log 2> gpurun_out/bench_adv.err; echo "== adversarial bench exit $? =="; tail -c 600 gpurun_out/bench_adv.log | cut -c1-400; tail -n 5 gpurun_out/bench_adv.err
fi
if [ "$1" == "prof" ]; then
  timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; cat gpurun_out/conv_bench.log
  CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
  timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
  timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -s 4400 -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
  echo "== launch list exit $? =="
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_wgrad|conv_tc_fwd|seg_loss_kernel|bn_bwd_apply" -s 200 -c 12 -o gpurun_out/prof_r01 $CMD > gpurun_out/ncu_full.log 2>&1
  echo "== full capture exit $? =="; tail -n 3 gpurun_out/ncu_full.log
fi
