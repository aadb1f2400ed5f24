#!/bin/bash
# round-end evidence: all GPU tests + smoke, micro-benchmarks, every bench workload, role traces (experiment build),
# ncu launch list of the bench command and one --set full capture of the dominant kernels (exported to CSV on the box)
R=${R:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in losses eval layers conv_tc bench_shapes bnfuse stages unet adversary; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -v "^E    +" gpurun_out/test_$f.log | tail -n 3
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -n 2 gpurun_out/smoke.log
timeout 120 tools/exp/umma_rate > gpurun_out/${R}_umma_rate.txt 2>&1; echo "== umma_rate exit $? =="
timeout 120 tools/exp/tmem_ld_rate > gpurun_out/${R}_tmem_ld_rate.txt 2>&1; echo "== tmem_ld_rate exit $? =="
timeout 600 python tools/timeline.py --dump --out gpurun_out/${R}_timeline.txt > gpurun_out/timeline.log 2>&1; echo "== timeline exit $? =="; head -n 5 gpurun_out/${R}_timeline.txt
timeout 600 python tools/trace_step.py > gpurun_out/${R}_trace_step.txt 2> gpurun_out/trace_step.err; echo "== trace_step exit $? =="; tail -n 9 gpurun_out/${R}_trace_step.txt
timeout 600 env UDA_B200_FUSE_BN_APPLY=0 python tools/trace_step.py > gpurun_out/${R}_trace_step_unfused.txt 2> gpurun_out/trace_step.err; echo "== trace_step (separate BatchNorm passes) exit $? =="; tail -n 7 gpurun_out/${R}_trace_step_unfused.txt
timeout 300 python tools/trace_conv.py layer1 layer2 layer3 layer4 l3.0 dec0.c1 dec2.c1 dec3.c1 > gpurun_out/${R}_trace_conv.txt 2>&1; echo "== trace exit $? =="
timeout 600 python tools/conv_bench.py > gpurun_out/${R}_conv_bench.txt 2>&1; echo "== conv_bench exit $? =="; tail -n 1 gpurun_out/${R}_conv_bench.txt
timeout 600 python tools/hbm_bench.py > gpurun_out/${R}_hbm_bench.txt 2>&1; echo "== hbm_bench exit $? =="
timeout 300 python tools/upconv_bench.py > gpurun_out/${R}_upconv_bench.txt 2>&1; echo "== upconv_bench exit $? =="
timeout 600 python tools/eval_bench.py > gpurun_out/${R}_eval_bench_4096.json 2> gpurun_out/eval_bench.err; echo "== eval_bench exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -n 3 gpurun_out/bench.err
timeout 900 python bench.py --steps 20 --warmup 5 --workload finetune --no-cpu-baseline > gpurun_out/${R}_bench_finetune.json 2> gpurun_out/bench_ft.err; echo "== bench finetune exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 --workload adversarial_grl --no-cpu-baseline > gpurun_out/${R}_bench_adversarial_grl.json 2> gpurun_out/bench_grl.err; echo "== bench grl exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 --workload adversarial --no-cpu-baseline > gpurun_out/${R}_bench_adversarial.json 2> gpurun_out/bench_adv.err; echo "== bench adversarial exit $? =="
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/bench_ref.err; echo "== bench reference exit $? =="
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv --log-file /tmp/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $? =="; wc -l /tmp/launches.csv
python tools/launch_shares.py /tmp/launches.csv gpurun_out/${R}_launch_shares.csv --step > /dev/null; head -n 12 gpurun_out/${R}_launch_shares.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv_tc_" -o /tmp/${R}_conv_full python tools/ncu_conv.py layer1 layer2 layer3 layer4 dec2.c1 > gpurun_out/ncu_full.log 2>&1
echo "== ncu full exit $? =="
ncu -i /tmp/${R}_conv_full.ncu-rep --page raw --csv > /tmp/raw.csv 2> gpurun_out/raw.err && python tools/ncu_select.py /tmp/raw.csv gpurun_out/${R}_ncu_full.csv
ncu -i /tmp/${R}_conv_full.ncu-rep --page source --csv > /tmp/source.csv 2> gpurun_out/source.err && python tools/ncu_source_top.py /tmp/source.csv 12 > gpurun_out/${R}_ncu_stall_sites.txt
ls -la gpurun_out | head -60; du -sm gpurun_out
