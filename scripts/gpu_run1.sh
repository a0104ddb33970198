#!/usr/bin/env bash
# round-2 GPU session 1: micro-benchmark, full GPU test-suite, the three bench workloads with breakdowns
mkdir -p gpurun_out
./scripts/micro/mma_m64.bin > gpurun_out/r02a_mma_m64.log 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x --timeout=900 > gpurun_out/r02a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02a_pytest.log
timeout 600 python bench.py --steps 50 --warmup 5 --profile-json gpurun_out/r02a_breakdown_train.json > gpurun_out/r02a_bench_train.json 2> gpurun_out/r02a_bench_train.err
timeout 600 python bench.py --workload wide_long --steps 20 --warmup 3 --profile-json gpurun_out/r02a_breakdown_wide.json > gpurun_out/r02a_bench_wide.json 2> gpurun_out/r02a_bench_wide.err
timeout 600 python bench.py --workload sample --steps 20 --warmup 3 --profile-json gpurun_out/r02a_breakdown_sample.json > gpurun_out/r02a_bench_sample.json 2> gpurun_out/r02a_bench_sample.err
tail -3 gpurun_out/r02a_pytest.log; cat gpurun_out/r02a_mma_m64.log | head -12
