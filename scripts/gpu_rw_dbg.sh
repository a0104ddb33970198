#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r04d}
for d in 0 8; do
  VAE_RW_DBG=$d timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-eager-baseline --profile-json gpurun_out/${TAG}_rwdbg${d}_train.json > /dev/null 2> gpurun_out/${TAG}_rwdbg${d}_train.err
done
ls gpurun_out/${TAG}_*
