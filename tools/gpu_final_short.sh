#!/bin/bash
# trimmed round-end evidence (no --set full capture, no micro-benchmarks): all GPU tests + smoke, role trace,
# transposed-conv micro-bench, every bench workload + the reference arm, ncu launch list of the bench command
R=${R:-r02}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
for f in losses eval layers conv_tc bench_shapes bnfuse stages unet adversary; do
  timeout 900 python -m pytest tests/test_gpu_$f.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_$f.log 2>&1
  echo "== test_gpu_$f exit $? =="; grep -v "^E    +" gpurun_out/test_$f.log | tail -n 2
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "== smoke exit $? =="; tail -n 2 gpurun_out/smoke.log
timeout 600 python tools/trace_step.py > gpurun_out/${R}_trace_step.txt 2> gpurun_out/trace_step.err; echo "== trace_step exit $? =="; tail -n 9 gpurun_out/${R}_trace_step.txt
timeout 300 python tools/upconv_bench.py > gpurun_out/${R}_upconv_bench.txt 2>&1; echo "== upconv_bench exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -n 3 gpurun_out/bench.err
timeout 900 python bench.py --steps 20 --warmup 5 --workload finetune --no-cpu-baseline > gpurun_out/${R}_bench_finetune.json 2> gpurun_out/bench_ft.err; echo "== bench finetune exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 --workload adversarial_grl --no-cpu-baseline > gpurun_out/${R}_bench_adversarial_grl.json 2> gpurun_out/bench_grl.err; echo "== bench grl exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 --workload adversarial --no-cpu-baseline > gpurun_out/${R}_bench_adversarial.json 2> gpurun_out/bench_adv.err; echo "== bench adversarial exit $? =="
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${R}_bench_reference.json 2> gpurun_out/bench_ref.err; echo "== bench reference exit $? =="
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-sub"
timeout 600 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 1500 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 12000 --csv --log-file /tmp/launches.csv $CMD > gpurun_out/ncu_launch.log 2>&1
echo "== ncu launches exit $? =="; wc -l /tmp/launches.csv
python tools/launch_shares.py /tmp/launches.csv gpurun_out/${R}_launch_shares.csv --step > /dev/null; head -n 14 gpurun_out/${R}_launch_shares.csv
for f in gpurun_out/${R}_bench*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), "roofline", {k:(d.get("roofline") or {}).get(k) for k in ("achieved","frac")})
PY
done
