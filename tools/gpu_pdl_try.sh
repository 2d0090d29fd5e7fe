#!/bin/bash
# A/B of the programmatic dependent launch chain: GPU tests with it on (default), short bench lines with it off / on,
# then the full default bench line.  usage (through gpurun): bash tools/gpu_pdl_try.sh TAG
TAG=${1:-r2g}
mkdir -p gpurun_out
bash tools/gpu_run_tests.sh
rc=$?
Q="--steps 10 --warmup 3 --no-train --no-cpu-baseline --no-parity-check --no-eager-baseline"
G4D_PDL=0 timeout -k 5 300 python bench.py $Q > gpurun_out/bench_pdl0_$TAG.json 2> gpurun_out/bench_pdl0_$TAG.err
G4D_PDL=1 timeout -k 5 300 python bench.py $Q > gpurun_out/bench_pdl1_$TAG.json 2> gpurun_out/bench_pdl1_$TAG.err
G4D_PDL=0 timeout -k 5 300 python bench.py $Q > gpurun_out/bench_pdl0b_$TAG.json 2>> gpurun_out/bench_pdl0_$TAG.err
G4D_PDL=1 timeout -k 5 300 python bench.py $Q > gpurun_out/bench_pdl1b_$TAG.json 2>> gpurun_out/bench_pdl1_$TAG.err
for f in pdl0 pdl1 pdl0b pdl1b; do python - <<P
import json
try:
    d = json.loads(open("gpurun_out/bench_${f}_$TAG.json").read().strip().splitlines()[-1])
    print("$f", round(d["value"], 1), "FPS  e2e", round(d["e2e"]["value"], 1), "stage", {k: round(v, 4) for k, v in d["stage_ms"].items() if v})
except Exception as e:
    print("$f", "failed", e)
P
done
if [ $rc -eq 0 ]; then
  timeout -k 5 900 python bench.py > gpurun_out/bench_n1_$TAG.json 2> gpurun_out/bench_n1_$TAG.err
  tail -c 600 gpurun_out/bench_n1_$TAG.json
fi
