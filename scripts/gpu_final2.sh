#!/usr/bin/env bash
# last check of the round: whole GPU suite, then the driver's default bench line and smoke()
mkdir -p gpurun_out
TAG=${1:-r04t}
bash scripts/gpu_tests.sh ${TAG}
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest.log | head
timeout 600 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/${TAG}_smoke.log 2>&1; tail -1 gpurun_out/${TAG}_smoke.log
timeout 900 python bench.py > gpurun_out/${TAG}_bench_line_train_n1.json 2> gpurun_out/${TAG}_bench_line_train_n1.err
python -c "
import json; b=json.loads(open('gpurun_out/${TAG}_bench_line_train_n1.json').read().strip().splitlines()[-1]); print('train', round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1), 'cpu', round(b['cpu_baseline']['value'],1), 'roofline', b['roofline']['kernel'], round(b['roofline']['frac'],3), b['clocks'])"
