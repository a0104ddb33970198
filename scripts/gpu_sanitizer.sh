#!/usr/bin/env bash
# ONE compute-sanitizer tool per call (B200_PROFILING.md) over the tcgen05 / TMA / mbarrier kernels' unit tests
mkdir -p gpurun_out
TOOL=${1:-memcheck}; TAG=${2:-r02u}
SEL='tests/test_umma_gpu.py tests/test_kernels_gpu.py::test_final_block_forward_backward tests/test_kernels_gpu.py::test_wide_latent_linear_kernels tests/test_kernels_gpu.py::test_skinny_head_gemms'
timeout 300 python -m pytest $SEL -m gpu -q -x --timeout=600 -k "not 64-64-64 and not B16" > gpurun_out/${TAG}_plain.log 2>&1 || { tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 1500 compute-sanitizer --tool $TOOL --error-exitcode 99 --print-limit 20 python -m pytest $SEL -m gpu -q -x --timeout=1400 -k "not 64-64-64 and not B16" > gpurun_out/${TAG}_${TOOL}.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/${TAG}_${TOOL}.log
grep -E "ERROR SUMMARY|passed|failed|rc=|Invalid|hazard" gpurun_out/${TAG}_${TOOL}.log | tail -8
