#!/bin/bash
# usage (through gpurun): bash tools/gpu_run_tests.sh [extra pytest args]   -> gpurun_out/pytest_gpu.txt
mkdir -p gpurun_out
rm -f gpurun_out/parity_fullsize.json
timeout -k 10 2400 python -m pytest tests -m gpu -q --timeout 900 -p no:cacheprovider "$@" > gpurun_out/pytest_gpu.txt 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -n 40 gpurun_out/pytest_gpu.txt
