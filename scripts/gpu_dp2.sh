#!/usr/bin/env bash
# 2 GPUs: the NCCL data-parallel parity test, then the train bench at N = 2
mkdir -p gpurun_out
TAG=${1:-r02o}
timeout 900 python -m pytest tests/test_dp_nccl_gpu.py -m gpu -q -rA --timeout=600 > gpurun_out/${TAG}_pytest_dp.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_dp.log
tail -4 gpurun_out/${TAG}_pytest_dp.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_train_n2.json 2> gpurun_out/${TAG}_bench_train_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload wide_long --steps 20 --warmup 3 --no-cpu-baseline --no-eager-baseline > gpurun_out/${TAG}_bench_wide_n2.json 2> gpurun_out/${TAG}_bench_wide_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 -m torch_vae_b200.train --batch-size 256 --image-size 128 --epochs 1 --n-rolls 8192 --log-interval 8 --models-dir gpurun_out/${TAG}_models > gpurun_out/${TAG}_launcher_n2.log 2>&1
tail -3 gpurun_out/${TAG}_launcher_n2.log
python - <<PY
import json
for w in ('train','wide'):
    try:
        b=json.loads(open('gpurun_out/${TAG}_bench_%s_n2.json' % w).read().strip().splitlines()[-1])
        print(w, 'n', b['n_gpus'], round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1))
    except Exception as e:
        print(w, 'failed', e)
PY
