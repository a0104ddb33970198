#!/usr/bin/env bash
# 2 GPUs, final commit: bench at N = 2 (train, wide / long) through the driver's torchrun command line, launcher under torchrun
mkdir -p gpurun_out
TAG=${1:-r04u}
run() {
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 --workload $2 --steps $3 --warmup $4 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_$2_n2.json 2> gpurun_out/${TAG}_bench_$2_n2.err
}
run 29551 train 100 5
run 29552 wide_long 20 3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 -m torch_vae_b200.train --batch-size 256 --image-size 128 --epochs 2 --n-rolls 8192 --log-interval 8 --models-dir gpurun_out/${TAG}_models > gpurun_out/${TAG}_launcher_n2.log 2>&1
grep -E "Epoch [0-9]|Error|error" gpurun_out/${TAG}_launcher_n2.log | tail -3
rm -rf gpurun_out/${TAG}_models
timeout 600 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q --timeout=500 2>&1 | tail -2
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/${TAG}_bench_*_n2.json')):
    try:
        b=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], 'n', b['n_gpus'], round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1))
    except Exception as e:
        print(f, 'failed', e, open(f.replace('.json','.err')).read()[-600:])
PY
