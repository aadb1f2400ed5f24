#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_bnfuse.py -q -m gpu --tb=short -p no:cacheprovider -k "maxpool" > gpurun_out/test_pool.log 2>&1; echo "== pool tests exit $? =="; tail -n 15 gpurun_out/test_pool.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_unet.py tests/test_gpu_stages.py -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/test_unet.log 2>&1; echo "== unet tests exit $? =="; tail -n 4 gpurun_out/test_unet.log | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/ab_pool1.json 2> gpurun_out/ab1.err; echo "== bench fused exit $? =="
timeout 600 env UDA_B200_FUSE_BN_POOL=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/ab_pool0.json 2> gpurun_out/ab0.err; echo "== bench unfused exit $? =="
for f in gpurun_out/ab_pool1.json gpurun_out/ab_pool0.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
kb=d.get("kernel_breakdown_ms_per_step") or {}
print(sys.argv[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, {k:v for k,v in kb.items() if "pool" in k or k in ("bn_apply_fused",)})
PY
done
