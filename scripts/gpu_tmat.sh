#!/usr/bin/env bash
# fused-phase transposed conv: unit tests + the inference test, then event timings at the wide and train shapes
mkdir -p gpurun_out
TAG=${1:-r03t}
timeout 900 python -m pytest tests/test_umma_gpu.py tests/test_step_parity_gpu.py -m gpu -q --timeout=600 -k "tmat or inference_epilogue or umma_tma" > gpurun_out/${TAG}_pytest_tmat.log 2>&1; echo "rc=$?" >> gpurun_out/${TAG}_pytest_tmat.log
grep -E "^(FAILED|ERROR)|passed|failed|rc=" gpurun_out/${TAG}_pytest_tmat.log | tail -12
REPS=5 python scripts/profile_tmat.py 2>&1 | tee gpurun_out/${TAG}_profile_tmat_plain.log
