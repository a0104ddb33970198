#!/usr/bin/env bash
mkdir -p gpurun_out
TAG=${1:-r04g}
for d in 0 1 0 1; do echo "== VAE_RW_DBG=$d"; VAE_RW_DBG=$d REPS=5 python scripts/profile_rw.py 2>&1 | grep -v "^$"; done | tee gpurun_out/${TAG}_rw_ab.log
