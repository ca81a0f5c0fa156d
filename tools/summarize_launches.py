"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum]` launch list by
kernel (our kernels by name + GEMM template arguments, library kernels by their leading identifier): launches, total time,
share of the step and — when the DRAM byte counters were collected — DRAM traffic and achieved GB/s per kernel family.
Usage: python tools/summarize_launches.py launches.csv [N] [--json out.json]"""
import csv
import json
import re
import sys
from collections import defaultdict


def load(path):
    lines = [l for l in open(path) if l.startswith('"')]
    per_launch = defaultdict(dict)       # launch id -> {metric: value, "name": kernel}
    for r in csv.DictReader(lines):
        n = r["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        if n.startswith("mv::"):
            m = re.match(r"mv::(\w+)(<[^>]*>)?", n)
            name = "mv::" + m.group(1) + ((m.group(2) or "") if "gemm" in m.group(1) else "")
        else:
            name = re.match(r"[\w:]+", n).group(0)
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        metric = r["Metric Name"]
        if metric == "gpu__time_duration.sum":
            v = v / 1e3 if unit.startswith("n") else (v * 1e3 if unit.startswith("m") else v)          # -> us
        else:
            v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)                # -> bytes
        per_launch[r["ID"]]["name"] = name
        per_launch[r["ID"]][metric] = v
    return per_launch


def main(path, top=40, out_json=None):
    per_launch = load(path)
    agg = defaultdict(lambda: [0, 0.0, 0.0])
    for d in per_launch.values():
        a = agg[d["name"]]
        a[0] += 1
        a[1] += d.get("gpu__time_duration.sum", 0.0)
        a[2] += d.get("dram__bytes_read.sum", 0.0) + d.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k.startswith("mv::"))
    have_bytes = any(a[2] > 0 for a in agg.values())
    print("total %.3f ms over %d launches; libmedvill_sm100 kernels %.3f ms (%.1f%%)" % (
        total / 1e3, sum(a[0] for a in agg.values()), ours / 1e3, 100 * ours / total))
    for name, (n, t, b) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        extra = "  %8.1f MB  %6.0f GB/s" % (b / 1e6, b / (t * 1e-6) / 1e9) if have_bytes else ""
        print("%8.3f ms %5.1f%% x%-4d%s  %s" % (t / 1e3, 100 * t / total, n, extra, name[:90]))
    if out_json:
        fam = {k: {"launches": n, "ms": t / 1e3, "dram_bytes": b} for k, (n, t, b) in agg.items()}
        gemm = [v for k, v in fam.items() if k.startswith("mv::gemm_tc05_kernel")]
        summary = {"step_kernel_ms": total / 1e3, "families": fam,
                   "gemm_tc05": {"launches": sum(v["launches"] for v in gemm), "ms": sum(v["ms"] for v in gemm),
                                 "dram_bytes": sum(v["dram_bytes"] for v in gemm)}}
        json.dump(summary, open(out_json, "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    oj = None
    if "--json" in args:
        i = args.index("--json")
        oj = args[i + 1]
        del args[i:i + 2]
    main(args[0], int(args[1]) if len(args) > 1 else 40, oj)
