#!/usr/bin/env bash
# ncu launch list of the train step (gpu__time_duration) + ncu --set full of the output-block kernels, each after the
# same command exited 0 without ncu
mkdir -p gpurun_out
TAG=${1:-r02i}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager-baseline --no-graph"
$CMD > gpurun_out/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_bench.log 2>&1
env B=512 REPS=2 python scripts/profile_final_bwd.py > gpurun_out/${TAG}_plain_final.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'final_' -s 4 -c 4 -o gpurun_out/${TAG}_output_block env B=512 REPS=2 python scripts/profile_final_bwd.py > gpurun_out/${TAG}_ncu_final.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_bench.log gpurun_out/${TAG}_ncu_final.log; ls -la gpurun_out | tail -8
