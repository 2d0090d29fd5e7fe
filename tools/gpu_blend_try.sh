#!/bin/bash
# A/B of the packed blend forward kernel: parity tests with it on, stage times with it on / off
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')" > /dev/null 2>&1
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q --timeout 600 -p no:cacheprovider -x > gpurun_out/pytest_blend.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_blend.txt
tail -n 6 gpurun_out/pytest_blend.txt
for m in 1 0; do
  G4D_BLEND_FWD_PACKED=$m timeout 300 python tools/profile_step.py --workload C3 --iters 6 --backward 0 --stage-times 1 > gpurun_out/prof_blend$m.txt 2>&1
  grep "^iter" gpurun_out/prof_blend$m.txt | tail -n 3
done
