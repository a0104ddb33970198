#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r03v}
for d in 0 1 2 4 6 8 9 15; do echo "== VAE_TMAT_DBG=$d"; VAE_TMAT_DBG=$d REPS=5 python scripts/profile_tmat.py 2>&1 | grep -v "^$"; done | tee gpurun_out/${TAG}_tmat_dbg.log
