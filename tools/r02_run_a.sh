# round-2 baseline run on one B200: new tests, bench with the new fields, copy ceiling at N=1, ncu baselines
set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_segments.py -x -q > gpurun_out/seg_test.log 2>&1; echo rc=$? >> gpurun_out/seg_test.log
python tools/copy_peak.py --gpus 1 --seconds 1.0 --out gpurun_out/copy_peak_n1.json > gpurun_out/copy_peak_n1.log 2>&1
python bench.py > gpurun_out/r02a_bench_c4.json 2> gpurun_out/r02a_bench_c4.err
python bench.py --workload c1 --cpu-frames 0 > gpurun_out/r02a_bench_c1.json 2> gpurun_out/r02a_bench_c1.err
ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_bilateral|k_upsharp" -s 40 -c 4 -o gpurun_out/prof_r02a python bench.py --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_r02a_b.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bicubic" -s 20 -c 2 -o gpurun_out/prof_r02a_c1 python bench.py --workload c1 --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_r02a_c1.log 2>&1
ls -la gpurun_out
tail -3 gpurun_out/seg_test.log
