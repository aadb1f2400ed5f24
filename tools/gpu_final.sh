#!/bin/bash
# round-end evidence: all GPU tests, micro-benchmarks, both bench workloads, ncu launch list + one full capture
mkdir -p gpurun_out
for f in losses eval layers conv_tc unet; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -v "^E    +" gpurun_out/test_$f.log | tail -n 4
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -n 2 gpurun_out/smoke.log
timeout 600 python tools/conv_bench.py > gpurun_out/conv_bench.log 2>&1; echo "== conv_bench exit $? =="; tail -n 1 gpurun_out/conv_bench.log
timeout 600 python tools/hbm_bench.py > gpurun_out/hbm_bench.log 2>&1; echo "== hbm_bench exit $? =="; cat gpurun_out/hbm_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -c 1500 gpurun_out/bench.log; tail -n 3 gpurun_out/bench.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload adversarial --no-cpu-baseline > gpurun_out/bench_adv.log 2> gpurun_out/bench_adv.err; echo "== adversarial bench exit $? =="; tail -c 300 gpurun_out/bench_adv.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 1300 -c 420 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $? =="; wc -l gpurun_out/launches.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_persist_kernel|conv_tc_wgrad_big_kernel|bn_bwd_apply_stream_kernel|seg_loss_stream_kernel" -s 150 -c 8 -o gpurun_out/prof_r01_final $CMD > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $? =="; tail -n 2 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
