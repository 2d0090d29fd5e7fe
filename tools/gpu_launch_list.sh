#!/bin/bash
# ncu launch list (gpu__time_duration per launch, --clock-control none) of a few forward and forward+backward C3 steps, plus a
# full-set capture of the forward kernels.  usage (through gpurun): bash tools/gpu_launch_list.sh TAG
TAG=${1:-r2}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_fwd_$TAG.csv \
  python tools/profile_step.py --workload C3 --iters 3 --backward 0 > gpurun_out/launches_fwd_$TAG.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bwd_$TAG.csv \
  python tools/profile_step.py --workload C3 --iters 3 --backward 1 > gpurun_out/launches_bwd_$TAG.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'deform_f16|deform_features|bin_|blend_forward' -c 12 -f -o gpurun_out/ncu_fwd_$TAG \
  python tools/profile_step.py --workload C3 --iters 2 --backward 0 > gpurun_out/ncu_fwd_$TAG.log 2>&1
ncu -i gpurun_out/ncu_fwd_$TAG.ncu-rep --page details --csv > gpurun_out/ncu_fwd_$TAG.details.csv 2>/dev/null
ncu -i gpurun_out/ncu_fwd_$TAG.ncu-rep --page raw --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor_subpipe_hmma.sum > gpurun_out/ncu_fwd_$TAG.raw.csv 2>/dev/null
tail -n 2 gpurun_out/ncu_fwd_$TAG.log
