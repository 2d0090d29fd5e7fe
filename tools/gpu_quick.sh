#!/bin/bash
# quick iteration loop: selected GPU tests (-k EXPR), the binning kernels under ncu, a short bench line.
# usage (through gpurun): bash tools/gpu_quick.sh TAG "pytest -k expression"
TAG=${1:-q}; K=${2:-"binning or indices or full_size or pdl or programmatic"}
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda()"
timeout -k 5 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider -k "$K" 2>&1 | tail -3
bash tools/gpu_bin_probe.sh $TAG | grep -E "time_duration|inst_executed.sum" | awk '{print $2, $3, $4}' | paste - - | head -6
timeout -k 5 300 python bench.py --steps 10 --no-cpu-baseline --no-parity-check --no-eager-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<P
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "FPS  e2e", round(d["e2e"]["value"], 1), "train", round(d["train_step"]["ms_per_step"], 3), {k: round(v, 4) for k, v in d["stage_ms"].items() if v})
P
