#!/bin/bash
# per-launch duration / instruction count / issue utilisation of the three binning kernels (ncu, 4 forward views of C3)
# usage (through gpurun): bash tools/gpu_bin_probe.sh TAG
TAG=${1:-probe}
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:'bin_' -c 24 --csv --log-file gpurun_out/binprobe_$TAG.csv \
  python tools/profile_step.py --workload C3 --iters 4 --backward 0 > gpurun_out/binprobe_$TAG.log 2>&1
python - <<P
import csv
rows = list(csv.reader(open("gpurun_out/binprobe_$TAG.csv")))
st = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
ix = {h: i for i, h in enumerate(rows[st])}
for r in rows[st + 1:]:
    if len(r) > ix["Metric Value"]:
        print(r[ix["ID"]], r[ix["Kernel Name"]][:18], r[ix["Metric Name"]][:40], r[ix["Metric Value"]])
P
