#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r04q}
N=${2:-2}
for cap in 25 1 0.25 25 1; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --bucket-cap-mb $cap > gpurun_out/${TAG}_cap${cap}.json 2> gpurun_out/${TAG}_cap${cap}.err
  python -c "
import json; b=json.loads(open('gpurun_out/${TAG}_cap${cap}.json').read().strip().splitlines()[-1]); print('cap $cap n', b['n_gpus'], round(b['value'],1), round(b['ms_per_step'],4))"
done
