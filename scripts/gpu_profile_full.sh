#!/usr/bin/env bash
# Run on the GPU box (via gpurun): launch list of the default bench + full ncu captures of the dominant kernels at the
# benchmark's own sizes.  Usage: bash scripts/gpu_profile_full.sh <tag>
TAG=${1:-r1}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
$CMD > gpurun_out/plain_full_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full_$TAG.csv $CMD > gpurun_out/ncu_launch_full_$TAG.log 2>&1
$CMD > gpurun_out/plain2_full_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_rk4_rollout -s 3 -c 1 -o gpurun_out/prof_rollout_full_$TAG $CMD > gpurun_out/ncu_rollout_full_$TAG.log 2>&1
$CMD > gpurun_out/plain3_full_$TAG.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_sens_fused -s 5 -c 1 -o gpurun_out/prof_sens_full_$TAG $CMD > gpurun_out/ncu_sens_full_$TAG.log 2>&1
ls -la gpurun_out/ | tail -12
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ekf_predict -s 2 -c 1 -o gpurun_out/prof_ekf_full_$TAG $CMD > gpurun_out/ncu_ekf_full_$TAG.log 2>&1
