#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_8gpu.log 2> gpurun_out/bench_8gpu.err; echo "== 8-GPU supervised exit $? =="; tail -c 1200 gpurun_out/bench_8gpu.log | cut -c1-1200; tail -n 3 gpurun_out/bench_8gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 10 --warmup 3 --workload adversarial > gpurun_out/bench_adv_8gpu.log 2> gpurun_out/bench_adv_8gpu.err; echo "== 8-GPU adversarial exit $? =="; tail -c 600 gpurun_out/bench_adv_8gpu.log | cut -c1-600; tail -n 3 gpurun_out/bench_adv_8gpu.err
