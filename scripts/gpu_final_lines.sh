#!/usr/bin/env bash
# the driver's own commands: default bench line (with cpu_baseline and the torch-eager leg) per workload, and the reference arm
mkdir -p gpurun_out
TAG=${1:-r04o}
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line_train_n1.json 2> gpurun_out/${TAG}_bench_line_train_n1.err
timeout 900 python bench.py --workload wide_long > gpurun_out/${TAG}_bench_line_wide_n1.json 2> gpurun_out/${TAG}_bench_line_wide_n1.err
timeout 900 python bench.py --workload sample > gpurun_out/${TAG}_bench_line_sample_n1.json 2> gpurun_out/${TAG}_bench_line_sample_n1.err
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_reference_arm_train.json 2> gpurun_out/${TAG}_reference_arm_train.err
python - <<PY
import json
for w in ('bench_line_train_n1','bench_line_wide_n1','bench_line_sample_n1','reference_arm_train'):
    try:
        b=json.loads(open('gpurun_out/${TAG}_%s.json' % w).read().strip().splitlines()[-1])
        print(w, round(b['value'],1), round(b.get('ms_per_step',0),4), 'e2e', round(b['e2e']['value'],1), 'cpu', b.get('cpu_baseline',{}).get('value'), 'eager', json.dumps(b.get('gpu_eager_baseline'))[:200], 'clocks', b.get('clocks'))
    except Exception as e:
        print(w, 'failed', e)
PY
