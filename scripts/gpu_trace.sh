#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r02j}
{
for m in 0 1; do SHAPE=512,32,32,64,128 python scripts/trace_conv2.py $m; done
for m in 0 1; do SHAPE=512,16,16,128,256 python scripts/trace_conv2.py $m; done
SHAPE=64,32,256,128,256 python scripts/trace_conv2.py 1
} > gpurun_out/${TAG}_trace_conv.log 2>&1
tail -5 gpurun_out/${TAG}_trace_conv.log
