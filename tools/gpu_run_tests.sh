#!/bin/bash
# usage (through gpurun): bash tools/gpu_run_tests.sh [extra pytest args]   -> gpurun_out/pytest_gpu.txt
# smoke() first under a short timeout: a kernel that deadlocks must cost two minutes of GPU budget, not forty.
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullsize.json
python -c "import torch; torch.zeros(1).cuda(); print('warm')" > gpurun_out/smoke.txt 2>&1   # a fresh box pages the image in: minutes, untimed
timeout -k 5 240 python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/smoke.txt 2>&1
rc=$?
echo "smoke rc=$rc" >> gpurun_out/smoke.txt
tail -n 3 gpurun_out/smoke.txt
if [ $rc -ne 0 ]; then echo "SMOKE FAILED: skipping the test suite"; tail -n 30 gpurun_out/smoke.txt; exit 1; fi
timeout -k 10 ${G4D_TEST_TIMEOUT:-1500} python -m pytest tests -m gpu -q --timeout 600 -p no:cacheprovider "$@" > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -n 25 gpurun_out/pytest_gpu.txt
