# round-2: single-warp mbarrier wait, bulk-copied bilateral tables, class LUTs, flattened expansion
set -x
T=${1:-r02c}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$? >> gpurun_out/${T}_gputest.log
tail -3 gpurun_out/${T}_gputest.log
python bench.py --cpu-frames 0 --no-c5-latency > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
python bench.py --workload c1 --cpu-frames 0 > gpurun_out/${T}_bench_c1.json 2> gpurun_out/${T}_bench_c1.err
python tools/show_bench.py gpurun_out/${T}_bench_*.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_c4.json"))
for k,v in d["roofline"]["kernels"].items(): print(k, v["avg_us"])
PY
ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_bilateral|k_upsharp" -s 40 -c 4 -o gpurun_out/prof_${T} python bench.py --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_${T}.log 2>&1
[ -x tools/pipebench ] && tools/pipebench > gpurun_out/pipebench.txt 2>&1
