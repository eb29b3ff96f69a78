"""Turns the ncu artefacts a gpurun call brought back into the small committed summaries under profiles/:
   python tools/ncu_summary.py <round-tag> <launches.csv> <prof.ncu-rep>
 - profiles/<tag>_launches.csv   : per-kernel launch count / average duration / share of the captured step
 - profiles/<tag>_ncu_summary.csv: one row per captured launch (duration, DRAM bytes, issue/pipe utilisation, stalls)
 - profiles/traffic.json         : dram__bytes_read.sum + dram__bytes_write.sum per launch, keyed like bench.py's kernels
 - profiles/kernel_counters.json : per kernel of ONE C4 frame (stats, bilateral_pre, upsharp, bilateral_post): warp instructions
                                   (smsp__inst_executed.sum), DRAM bytes, duration - bench.py's issue roofline reads it
"""
import collections
import csv
import json
import os
import subprocess
import sys

tag, launches_csv, rep = sys.argv[1:4]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)

rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    try:
        agg.setdefault(r[ki].split("(")[0], []).append(float(r[vi].replace(",", "")))
    except ValueError:
        pass
tot = sum(sum(v) for v in agg.values())
with open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow(["kernel", "launches", "avg_us", "share_of_captured_time"])
    for k, v in agg.items():
        w.writerow([k, len(v), round(sum(v) / len(v) / 1e3, 2), round(sum(v) / tot, 4)])

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
h = rr[0]
keys = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
keys = [k for k in keys if k in h]
units = rr[1]
traffic = {}
counters = collections.OrderedDict()
seen = collections.Counter()      # launches of the same kernel are told apart by their order in the captured frame
with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_summary.csv"), "w") as f:
    w = csv.writer(f)
    w.writerow([f"{k} [{units[h.index(k)]}]" if units[h.index(k)] else k for k in keys])
    for r in rr[2:]:
        w.writerow([r[h.index(k)][:60] for k in keys])
        name, grid = r[h.index("Kernel Name")], int(r[h.index("launch__grid_size")].replace(",", ""))
        ur, uw = units[h.index("dram__bytes_read.sum")], units[h.index("dram__bytes_write.sum")]
        scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        b = float(r[h.index("dram__bytes_read.sum")]) * scale.get(ur, 1) + float(r[h.index("dram__bytes_write.sum")]) * scale.get(uw, 1)
        base = name.split("(")[0]
        if "k_bilateral" in base:          # k_bilateral<0> (3-byte source) and k_bilateral<1> (32-bit source) are one kernel family
            base = "k_bilateral"
        traffic.setdefault(base + f"@grid{grid}#{seen[base]}", []).append(int(b))
        kind = {"k_stats": "stats", "k_bilateral": "bilateral_pre" if seen[base] == 0 else "bilateral_post"}.get(base)
        if kind is None and "k_upsharp" in base:
            kind = "upsharp"
        if kind and "smsp__inst_executed.sum" in h:
            counters[kind] = {"warp_inst": int(float(r[h.index("smsp__inst_executed.sum")])), "dram_bytes": int(b),
                              "us": round(float(r[h.index("gpu__time_duration.sum")]), 2), "grid": grid}
        seen[base] += 1
kj = os.path.join(ROOT, "profiles", "kernel_counters.json")
oldk = json.load(open(kj)) if os.path.exists(kj) else {}
if counters:
    oldk[tag] = counters
    json.dump(oldk, open(kj, "w"), indent=1, sort_keys=True)
tj = os.path.join(ROOT, "profiles", "traffic.json")
old = json.load(open(tj)) if os.path.exists(tj) else {}
old[tag] = {k: int(sum(v) / len(v)) for k, v in traffic.items()}
json.dump(old, open(tj, "w"), indent=1, sort_keys=True)
print(open(os.path.join(ROOT, "profiles", f"{tag}_launches.csv")).read())
print(json.dumps(old[tag], indent=1))
