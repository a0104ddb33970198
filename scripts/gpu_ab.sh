#!/usr/bin/env bash
# A/B of the two conv_umma_tma.cu pipeline switches on the mid / wide layers
mkdir -p gpurun_out
TAG=${1:-r02l}
for cfg in "0 0" "0 1" "1 0" "1 1"; do
  set -- $cfg
  echo "== VAE_TMA_DUAL=$1 VAE_TMA_MMA1T=$2"
  for sh in "512,32,32,64,128 0" "512,32,32,64,128 1" "512,16,16,128,256 0" "512,16,16,128,256 1" "64,32,256,128,256 1"; do
    set -- $cfg $sh
    VAE_TMA_DUAL=$1 VAE_TMA_MMA1T=$2 SHAPE=$3 python scripts/trace_conv2.py $4 | grep -E "^SHAPE|steady"
  done
done > gpurun_out/${TAG}_ab.log 2>&1
cat gpurun_out/${TAG}_ab.log
