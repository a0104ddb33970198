#!/usr/bin/env bash
# tests (no early stop) + the headline train workload with its per-kernel breakdown
mkdir -p gpurun_out
TAG=${1:-r02d}
bash scripts/gpu_tests.sh ${TAG}
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_train.json > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
python - <<PY
import json
b=json.loads(open('gpurun_out/${TAG}_bench_train.json').read())
print('train', b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], 'launches/step', b['gpu_launches_per_step'], b['roofline']['kernel'], b['roofline']['frac'])
PY
