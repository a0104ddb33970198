#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r02s}
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port"
A="-m torch_vae_b200.train --batch-size 256 --image-size 128 --epochs 2 --n-rolls 8192 --log-interval 8 --models-dir gpurun_out/${TAG}_models"
timeout 300 $L 29531 $A > gpurun_out/${TAG}_a.log 2>&1; echo "a rc=$?"
rm -rf gpurun_out/${TAG}_models
grep -E "^\[rank0\]:.*(frame #|c10d|Event|Stream|nccl)" gpurun_out/${TAG}_a.log | head -30
grep -E "Epoch|Saved" gpurun_out/${TAG}_a.log | tail -3
