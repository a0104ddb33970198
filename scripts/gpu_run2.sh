#!/usr/bin/env bash
# tests (no early stop) + the wide / long workload after the 512-channel split
mkdir -p gpurun_out
TAG=${1:-r02c}
bash scripts/gpu_tests.sh ${TAG}
timeout 600 python bench.py --workload wide_long --steps 20 --warmup 3 --no-cpu-baseline --profile-json gpurun_out/${TAG}_breakdown_wide.json > gpurun_out/${TAG}_bench_wide.json 2> gpurun_out/${TAG}_bench_wide.err
tail -c 600 gpurun_out/${TAG}_bench_wide.json
