#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r03g}
for cfg in "0 0" "2 2" "4 2" "1 3"; do
  set -- $cfg
  echo "== VAE_TMA_SPS=$1 VAE_TMA_MIN_STAGES=$2 (0 = default heuristic)"
  for sh in "512,32,32,64,128 0" "512,32,32,64,128 1" "512,16,16,128,256 0" "512,16,16,128,256 1" "64,32,256,128,256 1" "64,128,1024,64,64 1"; do
    set -- $cfg $sh
    VAE_TMA_SPS=$1 VAE_TMA_MIN_STAGES=$2 SHAPE=$3 python scripts/trace_conv2.py $4 | grep -E "^SHAPE|stages,"
  done
done > gpurun_out/${TAG}_sps.log 2>&1
cat gpurun_out/${TAG}_sps.log
