"""ncu launch-list CSVs (tools/gpu_launch_list.sh: --metrics gpu__time_duration.sum --clock-control none) -> the markdown table
committed under profiles/.

    python tools/summarize_launches.py gpurun_out/launches_fwd_TAG.csv gpurun_out/launches_bwd_TAG.csv profiles/TAG_launches_C3_fwd_bwd.md
"""
import collections
import csv
import sys


def table(path):
    rows = list(csv.reader(open(path)))
    st = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
    ix = {h: i for i, h in enumerate(rows[st])}
    acc = collections.OrderedDict()
    for r in rows[st + 1:]:
        if len(r) <= ix["Metric Value"] or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        acc.setdefault(r[ix["Kernel Name"]].split("(")[0][:70], []).append(us)
    total = sum(sum(v) for v in acc.values())
    out = ["| kernel | launches | mean us | share |", "|---|---|---|---|"]
    for k, v in sorted(acc.items(), key=lambda kv: -sum(kv[1])):
        out.append("| `%s` | %d | %.1f | %.1f %% |" % (k, len(v), sum(v) / len(v), 100.0 * sum(v) / total))
    return total, out


def main(fwd, bwd, md):
    with open(md, "w") as f:
        f.write("# ncu launch list, C3 (`ncu --metrics gpu__time_duration.sum --clock-control none`, tools/gpu_launch_list.sh)\n\n"
                "Per-launch times are cold-cache and serialised (and the dependent-launch chain of DESIGN.md 4.6 does not overlap under ncu): "
                "compare SHARES with the CUDA-event stage times of the bench line, not absolutes.\n")
        for title, path in (("3 fused forward steps (no grad)", fwd), ("3 fused forward + backward steps", bwd)):
            total, rows = table(path)
            f.write("\n## %s -- %.0f us of kernels\n\n" % (title, total) + "\n".join(rows) + "\n")
    print(open(md).read())


if __name__ == "__main__":
    main(*sys.argv[1:4])
