#!/usr/bin/env bash
# one 4-GPU box: smoke(), then train at N = 1 / 2 / 4 and wide / long at N = 4 (the driver's torchrun command line)
mkdir -p gpurun_out
TAG=${1:-r04p}
timeout 600 python -c 'import __graft_entry__ as g; g.smoke(); print("smoke ok")' > gpurun_out/${TAG}_smoke.log 2>&1; tail -2 gpurun_out/${TAG}_smoke.log
run() {  # run <n> <port> <workload> <steps> <warmup>
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $2 bench.py --gpus $1 --workload $3 --steps $4 --warmup $5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_$3_n$1.json 2> gpurun_out/${TAG}_bench_$3_n$1.err
}
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_train_n1.json 2> gpurun_out/${TAG}_bench_train_n1.err
run 2 29531 train 100 5
run 4 29532 train 100 5
run 4 29533 wide_long 20 3
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*_n*.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'n', b['n_gpus'], round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1))
    except Exception as e:
        print(f, 'failed', e)
PY
