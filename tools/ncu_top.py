"""Summarise an .ncu-rep: headline metrics + the SASS lines with the most stall samples, grouped by stall reason.
Usage: python tools/ncu_top.py file.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
want = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "launch__occupancy_limit", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_issue_stalled", "smsp__average_warps_issue_stalled"]
for i, h in enumerate(hdr):
    if any(h == w or (w.endswith("limit") and h.startswith(w)) or (w.startswith("smsp__average") and h.startswith(w)) for w in want):
        print("%-78s %s %s" % (h, vals[i], rows[1][i] if len(rows) > 2 else ""))
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]
data = [r for r in rows[2:] if len(r) == len(hdr)]
iS, iN = hdr.index("Source"), hdr.index("# Samples")
stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[iN] or 0) for r in data)
print("total samples", tot)
agg = {hdr[i]: sum(int(r[i] or 0) for r in data) for i in stall_cols}
print("stall reasons:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(1, tot)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
for r in sorted(data, key=lambda r: -int(r[iN] or 0))[:top]:
    reasons = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
    print("%6s %5.1f%%  %-70s %s" % (r[iN], 100 * int(r[iN]) / tot, r[iS][:70], " ".join("%s:%d" % (n, c) for c, n in reasons if c)))
