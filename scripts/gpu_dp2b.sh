#!/usr/bin/env bash
# 2 GPUs: the whole GPU suite (the NCCL parity tests run here), the launcher under torchrun, the train bench at N = 1
mkdir -p gpurun_out
TAG=${1:-r02r}
bash scripts/gpu_tests.sh ${TAG}
grep -E "^(FAILED|ERROR)" gpurun_out/${TAG}_pytest.log | head
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 -m torch_vae_b200.train --batch-size 256 --image-size 128 --epochs 2 --n-rolls 8192 --log-interval 8 --models-dir gpurun_out/${TAG}_models > gpurun_out/${TAG}_launcher_n2.log 2>&1
grep -E "Epoch|Saved|Error|error" gpurun_out/${TAG}_launcher_n2.log | tail -6
rm -rf gpurun_out/${TAG}_models
timeout 600 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_breakdown_train.json > gpurun_out/${TAG}_bench_train.json 2> gpurun_out/${TAG}_bench_train.err
python - <<PY
import json
b=json.loads(open('gpurun_out/${TAG}_bench_train.json').read())
print('train', round(b['value'],1), round(b['ms_per_step'],4), 'e2e', round(b['e2e']['value'],1), 'launches/step', b['gpu_launches_per_step'])
PY
