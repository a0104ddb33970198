#!/usr/bin/env bash
# kernel unit tests + step parity, then the train and wide benches with breakdowns
mkdir -p gpurun_out
TAG=${1:-r03b}
timeout 1200 python -m pytest tests/test_umma_gpu.py tests/test_kernels_gpu.py tests/test_step_parity_gpu.py tests/test_bf16_rigor_gpu.py -m gpu -q --timeout=900 > gpurun_out/${TAG}_pytest_sel.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_sel.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${TAG}_pytest_sel.log | tail -12
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_train.json > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
timeout 600 python bench.py --workload wide_long --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_wide.json > gpurun_out/${TAG}_bench_wide.json 2> gpurun_out/${TAG}_bench_wide.err
timeout 600 python bench.py --workload sample --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_sample.json > gpurun_out/${TAG}_bench_sample.json 2> gpurun_out/${TAG}_bench_sample.err
python - <<PY
import json
for w in ('train','wide','sample'):
    try:
        b=json.loads(open('gpurun_out/${TAG}_bench_%s.json' % w).read())
        print(w, round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1), 'launches/step', b['gpu_launches_per_step'], b['roofline']['kernel'], round(b['roofline']['frac'],3))
    except Exception as e:
        print(w, 'failed', e)
PY
