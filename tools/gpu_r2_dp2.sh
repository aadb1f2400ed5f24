#!/bin/bash
# the driver's multi-GPU command line (default settings) at N GPUs
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "== bench N=$N exit $? =="
python - "$N" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/r02_bench_{sys.argv[1]}gpu.json').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', round(d['e2e']['value'],1), d['config']['launch'][:100])
    for k in ('adversarial','adversarial_grl'):
        if k in d: print(' ', k, {q:round(d[k][q],2) for q in ('value','ms_per_step')}, 'e2e', round(d[k]['e2e']['value'],1), d[k]['config']['launch'][:80])
except Exception as e: print('bench parse failed', e)
PY
grep -v "^\s*$\|OMP_NUM\|\*\*\*\*" gpurun_out/bench_${N}gpu.err | tail -n 8
