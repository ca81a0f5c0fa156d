"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel (our kernels by name + GEMM template
arguments, library kernels by their leading identifier).  Usage: python tools/summarize_launches.py launches.csv [N]"""
import csv
import re
import sys
from collections import defaultdict


def main(path, top=34):
    lines = [l for l in open(path) if l.startswith('"')]
    agg = defaultdict(lambda: [0, 0.0])
    total = 0.0
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        n = r["Kernel Name"].replace("<unnamed>::", "").replace("void ", "")
        if n.startswith("mv::"):
            m = re.match(r"mv::(\w+)(<[^>]*>)?", n)
            name = "mv::" + m.group(1) + ((m.group(2) or "") if "gemm" in m.group(1) else "")
        else:
            name = re.match(r"[\w:]+", n).group(0)
        v = float(r["Metric Value"].replace(",", ""))
        v_us = v / 1e3 if r["Metric Unit"].startswith("n") else v
        agg[name][0] += 1
        agg[name][1] += v_us
        total += v_us
    ours = sum(t for k, (n, t) in agg.items() if k.startswith("mv::"))
    print("total %.3f ms over %d launches; libmedvill_sm100 kernels %.3f ms (%.1f%%)" % (
        total / 1e3, sum(a[0] for a in agg.values()), ours / 1e3, 100 * ours / total))
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
        print("%8.3f ms %5.1f%% x%-4d %s" % (t / 1e3, 100 * t / total, n, name[:100]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 34)
