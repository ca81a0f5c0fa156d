"""Per-source-line view of an ncu report captured with --import-source on: samples and executed instructions summed over the
SASS of each CUDA line (files under csrc/), top N lines.   python tools/ncu_lines.py report.ncu-rep [N]"""
import csv
import io
import subprocess
import sys


def main():
    rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    cur_file, hdr, out = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            i_s, i_i = hdr.index("# Samples"), hdr.index("Instructions Executed")
            continue
        if hdr is None or not r[0]:
            continue
        try:
            out.append((int(r[i_s]), cur_file, int(r[0]), int(r[i_i]), r[1].strip()[:100]))
        except (ValueError, IndexError):
            pass
    tot = sum(o[0] for o in out) or 1
    print("total samples %d, warp instructions %d" % (tot, sum(o[3] for o in out)))
    for smp, f, ln, ins, src in sorted(out, reverse=True)[:top]:
        print("%6d %5.1f%%  inst %10d  %s:%-4d %s" % (smp, 100.0 * smp / tot, ins, f, ln, src))


if __name__ == "__main__":
    main()
