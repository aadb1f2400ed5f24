#!/bin/bash
# refresh the bench lines after the roofline accounting fix (fused conv + BatchNorm launches charged to the conv family)
R=${R:-r02}
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${R}_bench.json 2> gpurun_out/bench.err; echo "== bench exit $? =="; tail -n 2 gpurun_out/bench.err
timeout 900 python bench.py --steps 20 --warmup 5 --workload finetune --no-cpu-baseline > gpurun_out/${R}_bench_finetune.json 2> gpurun_out/bench_ft.err; echo "== bench finetune exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 --workload adversarial_grl --no-cpu-baseline > gpurun_out/${R}_bench_adversarial_grl.json 2> gpurun_out/bench_grl.err; echo "== bench grl exit $? =="
timeout 900 python bench.py --steps 20 --warmup 5 --workload adversarial --no-cpu-baseline > gpurun_out/${R}_bench_adversarial.json 2> gpurun_out/bench_adv.err; echo "== bench adversarial exit $? =="
timeout 600 env UDA_B200_WGRAD_PRIORITY=1 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-sub > gpurun_out/ab_wgrad_priority.json 2> gpurun_out/ab.err; echo "== A/B wgrad priority exit $? =="
for f in gpurun_out/${R}_bench.json gpurun_out/${R}_bench_finetune.json gpurun_out/${R}_bench_adversarial_grl.json gpurun_out/${R}_bench_adversarial.json gpurun_out/ab_wgrad_priority.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], {k:d.get(k) for k in ("value","ms_per_step","gpu_launches")}, "e2e", (d.get("e2e") or {}).get("value"), "roofline", {k:(d.get("roofline") or {}).get(k) for k in ("achieved","frac","ms_per_step","fused_bn_launches_ms_per_step")})
PY
done
