#!/usr/bin/env bash
TAG=${1:-x}
CMD="python bench.py --traj 262144 --horizon 20 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_sens_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_sens -s 2 -c 2 -o gpurun_out/prof_sens_$TAG $CMD > gpurun_out/ncu_sens_$TAG.log 2>&1
