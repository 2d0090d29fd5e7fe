#!/bin/bash
# usage (through gpurun): bash tools/gpu_ncu.sh TAG KERNEL_REGEX [backward=0|1] [count]
#   -> gpurun_out/ncu_TAG.ncu-rep (+ .txt summary).  One GPU, never a bench number.
TAG=$1; RE=$2; BWD=${3:-0}; CNT=${4:-6}
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$RE -c $CNT -f -o gpurun_out/ncu_$TAG \
  python tools/profile_step.py --workload C3 --iters 3 --backward $BWD > gpurun_out/ncu_$TAG.log 2>&1
ncu -i gpurun_out/ncu_$TAG.ncu-rep --page details --csv > gpurun_out/ncu_$TAG.details.csv 2>/dev/null
tail -n 3 gpurun_out/ncu_$TAG.log
