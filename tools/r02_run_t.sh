# GPU parity suite + the default bench line
set -x
T=${1:-r02t}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$? >> gpurun_out/${T}_gputest.log
tail -3 gpurun_out/${T}_gputest.log
python bench.py --cpu-frames 0 --no-c5-latency > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
python tools/show_bench.py gpurun_out/${T}_bench_*.json
python - <<PY
import json
d=json.load(open("gpurun_out/${T}_bench_c4.json"))
for k,v in d["roofline"]["kernels"].items(): print(k, v["avg_us"])
PY
