"""Turn an `ncu --set full` report (exported with --page details --csv, tools/gpu_ncu.sh) into the markdown summary and the
traffic JSON committed under profiles/ (bench.py reads profiles/traffic.json for `roofline.traffic`).

    python tools/summarize_ncu.py gpurun_out/ncu_fwd.details.csv profiles/r2_ncu_full_C3_fwd.md profiles/traffic.json
"""
import collections
import csv
import json
import sys

WANT = ["Duration", "DRAM Throughput", "Memory Throughput", "L2 Cache Throughput", "Compute (SM) Throughput", "Registers Per Thread",
        "Achieved Occupancy", "Executed Ipc Active", "Issue Slots Busy", "Dynamic Shared Memory Per Block"]
STAGE = {"deform_features_kernel": "geom", "deform_tc_kernel": "geom", "deform_f16_kernel": "geom", "deform_kernel": "geom", "bin_sort_kernel": "binning",
         "bin_place_kernel": "binning", "bin_fix_kernel": "binning", "blend_forward_kernel": "blend"}


def main(src, md, tj):
    rows = list(csv.reader(open(src)))
    ix = {h: i for i, h in enumerate(rows[0])}
    per = collections.OrderedDict()
    for r in rows[1:]:
        if len(r) <= ix["Metric Value"]:
            continue
        key = (r[ix["ID"]], r[ix["Kernel Name"]].split("(")[0])
        per.setdefault(key, {})[r[ix["Metric Name"]]] = (r[ix["Metric Value"]], r[ix["Metric Unit"]])
    # dram bytes need the raw page; details give "DRAM Throughput" (%) and duration -> also accept explicit byte metrics if present
    agg = collections.OrderedDict()
    for (_, name), m in per.items():
        a = agg.setdefault(name, collections.defaultdict(list))
        for k, (v, u) in m.items():
            try:
                a[k + " [" + u + "]"].append(float(v.replace(",", "")))
            except ValueError:
                pass
    with open(md, "w") as f:
        f.write("# ncu --set full summary (%s)\n\nMean over the captured launches of each kernel; `--clock-control none`, cold caches, "
                "serialised: compare SHARES, not absolutes.\n\n" % src)
        cols = [c for c in ("Duration [us]", "Duration [ms]", "DRAM Throughput [%]", "Memory Throughput [%]", "L2 Cache Throughput [%]",
                            "Compute (SM) Throughput [%]", "Issue Slots Busy [%]", "Executed Ipc Active [inst/cycle]",
                            "Achieved Occupancy [%]", "Registers Per Thread [register/thread]")
                if any(c in a for a in agg.values())]
        f.write("| kernel | launches | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
        for name, a in agg.items():
            n = max(len(v) for v in a.values())
            f.write("| `%s` | %d | " % (name.replace("g4d::", ""), n) + " | ".join("%.2f" % (sum(a[c]) / len(a[c])) if c in a else "" for c in cols) + " |\n")
    print(open(md).read())
    if tj:
        # traffic: DRAM throughput fraction x measured HBM peak x duration is only an estimate; prefer dram__bytes metrics when
        # the raw page was exported next to the details file (same stem + .raw.csv)
        raw = src.replace(".details.csv", ".raw.csv")
        per_kernel = {}
        try:
            rr = list(csv.reader(open(raw)))
            h = {x: i for i, x in enumerate(rr[0])}
            rd = [k for k in h if k.startswith("dram__bytes_read.sum") and "per_second" not in k and "pct" not in k][0]
            wr = [k for k in h if k.startswith("dram__bytes_write.sum") and "per_second" not in k and "pct" not in k][0]
            units = rr[1]
            scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            acc = collections.defaultdict(list)
            for r in rr[2:]:
                name = r[h["Kernel Name"]].split("(")[0]
                b = float(r[h[rd]].replace(",", "")) * scale.get(units[h[rd]], 1.0) + float(r[h[wr]].replace(",", "")) * scale.get(units[h[wr]], 1.0)
                acc[name].append(b)
            per_kernel = {k: sum(v) / len(v) for k, v in acc.items()}
        except Exception as e:
            print("no raw page (%s): traffic.json not written" % e)
            return
        per_stage = collections.defaultdict(float)
        for k, b in per_kernel.items():
            for pat, st in STAGE.items():
                if pat in k:
                    per_stage[st] += b
        json.dump({"workload": "C3", "source": md + " (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum per launch)",
                   "bytes_per_launch": per_kernel, "per_stage": per_stage}, open(tj, "w"), indent=1)
        print(json.dumps(per_stage))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
