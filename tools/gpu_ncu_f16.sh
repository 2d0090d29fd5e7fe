mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:deform_f16_kernel -c 2 -f -o gpurun_out/ncu_f16 python tools/profile_step.py --workload C3 --iters 2 --backward 0 > gpurun_out/ncu_f16.log 2>&1
ncu -i gpurun_out/ncu_f16.ncu-rep --page details --csv > gpurun_out/ncu_f16.details.csv 2>/dev/null
tail -n 3 gpurun_out/ncu_f16.log
