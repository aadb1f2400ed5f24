#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_unet.py -q -m gpu --tb=short -p no:cacheprovider -k "graphed" > gpurun_out/test_graphed.log 2>&1; echo "== graphed tests exit $? =="; grep -v "^E    +" gpurun_out/test_graphed.log | tail -n 12
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --workload adversarial > gpurun_out/bench_adv_2gpu.log 2> gpurun_out/bench_adv_2gpu.err; echo "== 2-GPU adversarial exit $? =="; tail -c 2500 gpurun_out/bench_adv_2gpu.log | cut -c1-700; tail -n 5 gpurun_out/bench_adv_2gpu.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu.log 2> gpurun_out/bench_2gpu.err; echo "== 2-GPU supervised exit $? =="; tail -c 2500 gpurun_out/bench_2gpu.log | cut -c1-300; tail -n 3 gpurun_out/bench_2gpu.err
