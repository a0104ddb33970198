#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r03z}
for d in 0 1 0 1; do echo "== VAE_TMAT_SPLIT=$d"; VAE_TMAT_SPLIT=$d REPS=5 python scripts/profile_tmat.py 2>&1 | grep -v "^$"; done | tee gpurun_out/${TAG}_tmat_split_ab.log
for d in 0 1; do VAE_TMAT_SPLIT=$d timeout 600 python -m pytest tests/test_umma_gpu.py tests/test_step_parity_gpu.py -m gpu -q --timeout=600 -k "tmat or inference_epilogue or umma_tma" 2>&1 | tail -2; done
