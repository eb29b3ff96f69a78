ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,regex:smsp__average_warps_issue_stalled_.*_per_issue_active.ratio,sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_varsharp --launch-skip 6 -c 2 --csv --log-file gpurun_out/vs_metrics.csv python bench.py --steps 1 --warmup 1 --frames 8 --cpu-frames 0 --no-e2e --adaptive-sigma > gpurun_out/vs_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/vs_metrics.csv')) if len(r)>10]
h=rows[0]; mi=h.index('Metric Name'); vi=h.index('Metric Value'); ii=h.index('ID')
for r in rows[1:]:
    if r[ii]==rows[1][ii]: print(r[mi], r[vi])
PY
