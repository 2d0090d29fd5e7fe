#!/bin/bash
# compute-sanitizer memcheck + racecheck of the smoke step (tiny fused forward + backward through render(), tensor-core kernels)
# usage (through gpurun): bash tools/gpu_sanitizer.sh TAG
TAG=${1:-r2f}
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()" > /dev/null 2>&1
for tool in memcheck racecheck; do
  timeout -k 5 ${G4D_SAN_TIMEOUT:-420} compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitizer_${tool}_$TAG.txt 2>&1
  echo "rc=$?" >> gpurun_out/sanitizer_${tool}_$TAG.txt
  grep -E "ERROR SUMMARY|smoke:|rc=" gpurun_out/sanitizer_${tool}_$TAG.txt | tail -n 3
done
