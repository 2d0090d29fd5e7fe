#!/bin/bash
# first contact of a new tensor-core kernel with the GPU: smoke under a short timeout, the focused parity tests, cycle counters
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')" > gpurun_out/smoke.txt 2>&1
timeout -k 5 200 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.txt 2>&1
rc=$?
echo "smoke rc=$rc" >> gpurun_out/smoke.txt
tail -n 4 gpurun_out/smoke.txt
if [ $rc -ne 0 ]; then echo "SMOKE FAILED"; tail -n 30 gpurun_out/smoke.txt; exit 1; fi
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 300 -p no:cacheprovider -x -k "tensor_core or fused or deform" > gpurun_out/pytest_f16.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_f16.txt
tail -n 15 gpurun_out/pytest_f16.txt
for m in 2 1; do
  timeout 300 python tools/profile_step.py --workload C3 --iters 4 --backward 0 --stage-times 1 --tc-debug 1 --tc-mode $m > gpurun_out/prof_mode$m.txt 2>&1
  tail -n 6 gpurun_out/prof_mode$m.txt
done
