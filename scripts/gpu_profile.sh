#!/usr/bin/env bash
# Run on the GPU box (via gpurun): launch list + full ncu capture of the rollout and sensitivity kernels.
# Usage: bash scripts/gpu_profile.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --traj 262144 --horizon 100 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_launch_$TAG.log 2>&1
$CMD > gpurun_out/plain2_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_rk4_rollout -s 1 -c 1 -o gpurun_out/prof_rollout_$TAG $CMD > gpurun_out/ncu_rollout_$TAG.log 2>&1
$CMD > gpurun_out/plain3_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_sens -s 2 -c 2 -o gpurun_out/prof_sens_$TAG $CMD > gpurun_out/ncu_sens_$TAG.log 2>&1
ls -la gpurun_out/
