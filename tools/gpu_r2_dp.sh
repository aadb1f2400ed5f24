#!/bin/bash
# data-parallel check: N GPUs (default 2) — supervised line with both adversarial sub-records, NCCL captured in the graphs;
# A/B against the flat all-reduce after the replay
N=${N:-2}
mkdir -p gpurun_out
run() { # name, extra args
  local name=$1; shift
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline "$@" > gpurun_out/bench_${name}_${N}gpu.log 2> gpurun_out/bench_${name}_${N}gpu.err
  echo "== bench $name N=$N exit $? =="
  python - "$name" "$N" <<'PY'
import json, sys
try:
    d=json.loads(open(f'gpurun_out/bench_{sys.argv[1]}_{sys.argv[2]}gpu.log').read().strip().splitlines()[-1])
    print({k:d[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', round(d['e2e']['value'],1), d['config']['launch'][:90])
    for k in ('adversarial','adversarial_grl'):
        if k in d: print(' ', k, {q:round(d[k][q],2) for q in ('value','ms_per_step')}, 'e2e', round(d[k]['e2e']['value'],1))
except Exception as e: print('bench parse failed', e)
PY
  grep -v "^\s*$" gpurun_out/bench_${name}_${N}gpu.err | tail -n 6
}
run ingraph
run flat --nccl-outside-graph --no-sub
