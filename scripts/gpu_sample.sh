#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r04j}
timeout 600 python -m pytest tests/test_step_parity_gpu.py tests/test_kernels_gpu.py -m gpu -q --timeout=600 -k "sample or final" -s 2>&1 | grep -E "sample_to_host|passed|failed|Error" | tail -5
for c in 512 1024 2048; do timeout 300 python bench.py --workload sample --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --sample-chunk $c > gpurun_out/${TAG}_bench_sample_c$c.json 2> gpurun_out/${TAG}_bench_sample_c$c.err; python -c "
import json; b=json.loads(open('gpurun_out/${TAG}_bench_sample_c$c.json').read()); print('chunk $c', round(b['value']), round(b['ms_per_step'],3), 'e2e', round(b['e2e']['value']), round(b['e2e']['ms_per_step'],3))"; done
timeout 600 python bench.py --workload wide_long --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_wide.json > gpurun_out/${TAG}_bench_wide.json 2> gpurun_out/${TAG}_bench_wide.err; python -c "
import json; b=json.loads(open('gpurun_out/${TAG}_bench_wide.json').read()); print('wide', round(b['value'],1), round(b['ms_per_step'],3), b['roofline']['kernel'], round(b['roofline']['avg_launch_us'],1))"
