#!/usr/bin/env bash
# wide / long workload on N GPUs with 0 / 8 / 16 SMs left to the NCCL kernels
mkdir -p gpurun_out
TAG=${1:-r03q}; N=${2:-4}
port=29540
for k in 0 8 16; do
  port=$((port+1))
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --workload wide_long --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline --reserve-sms $k > gpurun_out/${TAG}_wide_n${N}_reserve$k.json 2> gpurun_out/${TAG}_wide_n${N}_reserve$k.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_wide_n*_reserve*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1]); print(f.split('/')[-1], round(b['value'],1), round(b['ms_per_step'],4))
    except Exception as e: print(f, 'failed', e)
PY
