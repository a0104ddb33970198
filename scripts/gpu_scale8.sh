#!/usr/bin/env bash
# scaling lines on one 8-GPU box: train at N = 8 and 4, wide / long at N = 8, the DP parity test, NCCL_DEBUG summary
mkdir -p gpurun_out
TAG=${1:-r02q}
run() {  # run <n> <port> <workload> <steps> <warmup>
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --workload $3 --steps $4 --warmup $5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_$3_n$1.json 2> gpurun_out/${TAG}_bench_$3_n$1.err
}
run 8 29521 train 100 5
run 8 29522 wide_long 20 3
run 4 29523 train 100 5
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_train_n1.json 2> gpurun_out/${TAG}_bench_train_n1.err
NCCL_DEBUG=INFO timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 8 --steps 5 --warmup 3 --no-cpu-baseline --no-eager-baseline 2>&1 | grep -E "NVLS|Channel|Connected|NET|P2P|nRanks" | sort | uniq -c | sort -rn | head -20 > gpurun_out/${TAG}_nccl_info.log
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*_n*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'n', b['n_gpus'], round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1))
    except Exception as e:
        print(f, 'failed', e)
PY
