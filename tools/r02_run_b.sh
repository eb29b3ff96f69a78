# round-2: integer-ratio bicubic kernel - parity, bench C1/C5, ncu
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_gputest.log 2>&1; echo rc=$? >> gpurun_out/r02b_gputest.log
python bench.py --workload c1 --cpu-frames 0 > gpurun_out/r02b_bench_c1.json 2> gpurun_out/r02b_bench_c1.err
python bench.py --workload c5 --cpu-frames 0 > gpurun_out/r02b_bench_c5.json 2> gpurun_out/r02b_bench_c5.err
ncu --set full --clock-control none --import-source on -k regex:"k_bicubic" -s 20 -c 2 -o gpurun_out/prof_r02b_c1 python bench.py --workload c1 --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_r02b_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bicubic" -s 20 -c 2 -o gpurun_out/prof_r02b_c5 python bench.py --workload c5 --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_r02b_c5.log 2>&1
tail -3 gpurun_out/r02b_gputest.log
python tools/show_bench.py gpurun_out/r02b_bench_*.json
