#!/usr/bin/env bash
# output-block tests, then the train bench with the deep-ring pass 1 off / on
mkdir -p gpurun_out
TAG=${1:-r02w}
timeout 900 python -m pytest tests/test_kernels_gpu.py::test_final_block_forward_backward tests/test_bf16_rigor_gpu.py tests/test_step_parity_gpu.py -m gpu -q --timeout=600 > gpurun_out/${TAG}_pytest_sel.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_sel.log
tail -3 gpurun_out/${TAG}_pytest_sel.log
for d in 0; do
  VAE_FINAL_DEEP=$d timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_train_deep$d.json > gpurun_out/${TAG}_bench_train_deep$d.json 2> gpurun_out/${TAG}_bench_train_deep$d.err
done
python - <<PY
import json
for d in (0,):
    b=json.loads(open('gpurun_out/${TAG}_bench_train_deep%d.json' % d).read())
    k=json.load(open('gpurun_out/${TAG}_breakdown_train_deep%d.json' % d))
    p1=[r for r in k['kernels'] if r['name']=='vae_final_bwd_reduce_tc'][0]
    print('deep',d, round(b['value'],1), round(b['ms_per_step'],4), 'pass1 us', round(p1['ms_per_step']*1000,1), b['roofline']['kernel'], round(b['roofline']['frac'],3))
PY
