#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r02n}
for sh in "512,32,32,64,128 0" "512,32,32,64,128 1" "512,16,16,128,256 0" "512,16,16,128,256 1" "64,32,256,128,256 1"; do
  set -- $sh
  SHAPE=$1 python scripts/trace_conv2.py $2 | grep -E "^SHAPE|steady"
done > gpurun_out/${TAG}_ab.log 2>&1
cat gpurun_out/${TAG}_ab.log; nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv
