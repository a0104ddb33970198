#!/usr/bin/env bash
# plain run, then ncu --set full of the fused-phase transposed conv kernel
mkdir -p gpurun_out
TAG=${1:-r03s}
REPS=5 python scripts/profile_tmat.py > gpurun_out/${TAG}_profile_tmat_plain.log 2>&1 &&
ncu --set full --import-source on --clock-control none -k regex:'umma_tmat_conv_kernel' -c 4 \
    -o gpurun_out/${TAG}_tmat env REPS=1 python scripts/profile_tmat.py > gpurun_out/${TAG}_profile_tmat_ncu.log 2>&1
cat gpurun_out/${TAG}_profile_tmat_plain.log; tail -3 gpurun_out/${TAG}_profile_tmat_ncu.log; ls -la gpurun_out/*.ncu-rep
