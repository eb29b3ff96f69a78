python bench.py > gpurun_out/r01f_bench_c4.json 2> gpurun_out/r01f_bench_c4.err
python bench.py --workload c2 --cpu-frames 0 > gpurun_out/r01f_bench_c2.json 2> /dev/null
python bench.py --workload c3 --cpu-frames 0 > gpurun_out/r01f_bench_c3.json 2> /dev/null
python bench.py --bilateral-mode selective --frames 64 --cpu-frames 6 > gpurun_out/r01f_bench_c4_selective.json 2> gpurun_out/r01f_sel.err
python bench.py --bilateral-mode multiscale --frames 64 --cpu-frames 6 > gpurun_out/r01f_bench_c4_multiscale.json 2> gpurun_out/r01f_ms.err
python bench.py --adaptive-sigma --frames 64 --cpu-frames 6 > gpurun_out/r01f_bench_c4_adaptive_sigma.json 2> gpurun_out/r01f_as.err
python bench.py --bilateral-mode multiscale --adaptive-sigma --flow --frames 32 --cpu-frames 3 > gpurun_out/r01f_bench_c4_reference_defaults.json 2> gpurun_out/r01f_rd.err
python bench.py --steps 2 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e > gpurun_out/plain_r01f.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_stats|k_bilateral|k_upsharp|k_blend|k_scene" -s 20 -c 200 --csv --log-file gpurun_out/launches_r01f.csv python bench.py --steps 2 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e > gpurun_out/ncu_r01f_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_bilateral|k_upsharp" -s 40 -c 4 -o gpurun_out/prof_r01f python bench.py --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_r01f_b.log 2>&1
