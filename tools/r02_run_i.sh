set -x
ncu --set full --clock-control none --import-source on -k regex:"k_bicubic" -s 20 -c 1 -o gpurun_out/prof_r02i_c1 python bench.py --workload c1 --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_r02i_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_bilateral|k_upsharp" -s 40 -c 4 -o gpurun_out/prof_r02i python bench.py --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_r02i.log 2>&1
AB_ARGS="--workload c1" bash tools/ab.sh r02i_c1 base
ls -la gpurun_out/*.ncu-rep
