#!/bin/bash
TAG=${1:-r2h}
mkdir -p gpurun_out
python -c "import torch; torch.zeros(1).cuda(); print('warm')"
timeout -k 5 600 python -m pytest tests/test_gpu_train_dp.py tests/test_gpu_parity.py -m gpu -q -x -p no:cacheprovider -k "train or ranks or adam or no_sync or batched or fused_render_forward" 2>&1 | tail -4
timeout -k 5 300 python tools/train_dp_profile.py > gpurun_out/train_dp_profile_$TAG.txt 2>&1; head -60 gpurun_out/train_dp_profile_$TAG.txt | cut -c1-200
timeout -k 5 300 python bench.py --steps 10 --no-cpu-baseline --no-parity-check --no-eager-baseline > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err
python - <<P
import json
d = json.loads(open("gpurun_out/bench_$TAG.json").read().strip().splitlines()[-1])
print(round(d["value"], 1), "FPS  e2e", round(d["e2e"]["value"], 1), "train", d["train_step"]["ms_per_step"], d["train_step"]["step_ms"])
P
