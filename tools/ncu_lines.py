"""Per-source-line instruction / stall-sample histogram of one kernel from an ncu report taken with
--import-source on (kernels compiled with -lineinfo).

    python tools/ncu_lines.py <prof.ncu-rep> <kernel-regex> [min_pct] [launch-id]

Prints file:line, share of warp instructions executed, share of stall samples, the source text.
"""
import csv
import subprocess
import sys

rep, rx = sys.argv[1], sys.argv[2]
min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + rx]
if len(sys.argv) > 4:
    cmd += ["--launch-skip", sys.argv[4], "--launch-count", "1"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
cur_file = None
hdr = None
lines = []   # (file, line, src, inst, samples)
seen_kernel = 0
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or r[0] == "" or not r[0].isdigit():
        continue
    try:
        lines.append((cur_file, int(r[0]), r[1], int(r[ii]), int(r[si])))
    except (ValueError, IndexError):
        pass
ti = sum(l[3] for l in lines) or 1
ts = sum(l[4] for l in lines) or 1
print(f"total warp instructions {ti}, samples {ts}")
for f, ln, src, inst, smp in lines:
    if 100.0 * inst / ti >= min_pct or 100.0 * smp / ts >= min_pct:
        print(f"{f}:{ln:<4d} inst {100.0 * inst / ti:5.1f}%  samples {100.0 * smp / ts:5.1f}%  {src.strip()[:110]}")
