#!/bin/bash
# round-end evidence at HEAD: full GPU suite, the driver's bench command, ncu launch lists + full capture of the forward kernels.
# usage (through gpurun): bash tools/gpu_final.sh TAG
TAG=${1:-final}
mkdir -p gpurun_out
bash tools/gpu_run_tests.sh > gpurun_out/tests_$TAG.log 2>&1; tail -n 3 gpurun_out/tests_$TAG.log
cp gpurun_out/pytest_gpu.txt gpurun_out/pytest_gpu_$TAG.txt
timeout -k 5 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
python - <<P
import json
d = json.loads(open("gpurun_out/bench_n1_$TAG.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "FPS  e2e", round(d["e2e"]["value"], 1), "train", round(d["train_step"]["ms_per_step"], 3), {k: round(v, 4) for k, v in d["stage_ms"].items() if v})
P
bash tools/gpu_launch_list.sh $TAG > /dev/null 2>&1
ls gpurun_out | grep $TAG | head -20
