"""Aggregate an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
total = 0.0
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v_us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    agg[name][0] += 1
    agg[name][1] += v_us
    total += v_us
print("total %.3f ms over %d launches" % (total / 1e3, sum(a[0] for a in agg.values())))
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print("%8.3f ms  %5.1f%%  x%-4d  %s" % (t / 1e3, 100 * t / total, n, name[:90]))
