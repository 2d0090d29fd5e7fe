#!/bin/bash
mkdir -p gpurun_out; : > gpurun_out/triage.txt
for args in ${TRIAGE_STAGES:-"raster_bwd:2000" "deform_bwd64:2000" "deform_bwd128:2000" "fused_bwd:2000" "fused_bwd_debug:2000" "smoke:0"}; do
  args=${args/:/ }
  timeout -k 2 ${TRIAGE_TIMEOUT:-50} python tools/gpu_triage.py $args >> gpurun_out/triage.txt 2>&1
  echo "[$args] rc=$?" >> gpurun_out/triage.txt
done
cat gpurun_out/triage.txt | grep -v Warning | tail -n 40
