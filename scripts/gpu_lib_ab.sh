#!/usr/bin/env bash
# A/B of two builds of the library on one box: bench breakdowns (train, sample) with each
mkdir -p gpurun_out
TAG=${1:-r04h}
ALT=/root/repo/torch-vae_b200/libvae_b200_generic.so
for v in main alt main alt; do
  if [ $v = alt ]; then export VAE_B200_LIB=$ALT; else unset VAE_B200_LIB; fi
  timeout 300 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_${v}_train.json > gpurun_out/${TAG}_${v}_train.line 2> /dev/null
  timeout 300 python bench.py --workload sample --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_${v}_sample.json > gpurun_out/${TAG}_${v}_sample.line 2> /dev/null
  python - <<PY
import json
for w in ('train','sample'):
    b=json.loads(open('gpurun_out/${TAG}_${v}_%s.line' % w).read()); k=json.load(open('gpurun_out/${TAG}_${v}_%s.json' % w))['kernels']
    top=' '.join('%s%s=%.0f' % (r['name'][4:], str(r['args'][:7]).replace(' ',''), r['ms_per_step']*1e3) for r in k[:9])
    print('$v', w, round(b['ms_per_step'],4), top)
PY
done
