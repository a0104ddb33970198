#!/usr/bin/env bash
# the three bench workloads with breakdowns (no tests), plus the train step without the gradient sink
mkdir -p gpurun_out
TAG=${1:-r04m}
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_train.json > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --no-grad-sink > gpurun_out/${TAG}_bench_train_nosink.json 2> gpurun_out/${TAG}_bench_train_nosink.err
timeout 600 python bench.py --workload wide_long --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_wide.json > gpurun_out/${TAG}_bench_wide.json 2> gpurun_out/${TAG}_bench_wide.err
timeout 600 python bench.py --workload sample --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_sample.json > gpurun_out/${TAG}_bench_sample.json 2> gpurun_out/${TAG}_bench_sample.err
python - <<PY
import json
for w in ('train','train_nosink','wide','sample'):
    try:
        b=json.loads(open('gpurun_out/${TAG}_bench_%s.json' % w).read())
        print(w, round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1), 'launches/step', b['gpu_launches_per_step'], b['roofline']['kernel'], round(b['roofline']['frac'],3))
    except Exception as e:
        print(w, 'failed', e)
PY
