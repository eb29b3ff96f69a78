"""One CSV row per captured launch of one or more ncu reports (the columns DESIGN.md quotes):
   python tools/ncu_rows.py out.csv rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import subprocess
import sys

KEYS = ["Kernel Name", "launch__grid_size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
out, reps = sys.argv[1], sys.argv[2:]
with open(out, "w", newline="") as f:
    w = None
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        h, units = rows[0], rows[1]
        idx = [h.index(k) for k in KEYS if k in h]
        if w is None:
            w = csv.writer(f)
            w.writerow(["capture"] + [f"{h[i]} [{units[i]}]" if units[i] else h[i] for i in idx])
        for r in rows[2:]:
            w.writerow([rep.split("/")[-1].replace(".ncu-rep", "")] + [r[i][:60] for i in idx])
