#!/usr/bin/env bash
# GPU tests, then ncu --set full of the wide-configuration conv kernels (after the same command exited 0 without ncu)
mkdir -p gpurun_out
TAG=${1:-r02g}
bash scripts/gpu_tests.sh ${TAG}
REPS=2 python scripts/profile_wide.py > gpurun_out/${TAG}_profile_wide_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:'umma_tma_conv_kernel|umma_wgrad_tma_kernel' -s 10 -c 10 \
    -o gpurun_out/${TAG}_wide_convs env REPS=2 python scripts/profile_wide.py > gpurun_out/${TAG}_profile_wide_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_profile_wide_plain.log; tail -3 gpurun_out/${TAG}_profile_wide_ncu.log; ls -la gpurun_out/*.ncu-rep
