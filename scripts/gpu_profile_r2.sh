#!/usr/bin/env bash
# Run on the GPU box (via gpurun): launch list of the default bench + full ncu captures of the dominant kernels at the
# benchmark's own sizes.  Usage: bash scripts/gpu_profile_r2.sh <tag> [kernels...]   (kernels: rollout idsweep sens sensroll ekf colloc collocsp)
TAG=${1:-r2}; shift
KS=${@:-rollout idsweep sens sensroll ekf}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_full_$TAG.log 2>&1 || { echo "plain bench failed"; tail -5 gpurun_out/plain_full_$TAG.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_full_$TAG.csv $CMD > gpurun_out/ncu_launch_full_$TAG.log 2>&1
cap() {   # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -o gpurun_out/prof_$1_$TAG -f $CMD > gpurun_out/ncu_$1_$TAG.log 2>&1
}
for k in $KS; do
  case $k in
    rollout)  cap rollout 'k_rk4_rollout' 4 ;;               # ncu matches the base name: launches 1-5 are <1,0,0> (config 2: 3 warm-up + 2 timed)
    idsweep)  cap idsweep 'k_rk4_rollout' 9 ;;               # launch 6 = the B = 1 measurement log, 7-11 = <2,0,1> (config 5)
    sensroll) cap sensroll 'k_sens_fused' 1 ;;               # first launches = the 1 M x 10 rollout
    sens)     cap sens 'k_sens_fused' 8 ;;                   # then the single steps
    ekf)      cap ekf 'k_ekf_predict' 2 ;;
    colloc)   cap colloc 'k_colloc_eval' 2 ;;
    collocsp) cap collocsp 'k_colloc_eval' 10 ;;             # launches 0-7 = dense blocks (3 warm-up + 5 timed), then the sparse format
  esac
done
ls -la gpurun_out/ | grep $TAG
