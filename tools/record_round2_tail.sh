# after the last kernel change of the round (PRE -> k_upsharp lane buffer): the headline records again, tag r02zz
set -x
T=${1:-r02zz}
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/${T}_smoke.log 2>&1; echo rc=$? >> gpurun_out/${T}_smoke.log; tail -2 gpurun_out/${T}_smoke.log
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$? >> gpurun_out/${T}_gputest.log; tail -2 gpurun_out/${T}_gputest.log
python bench.py > gpurun_out/${T}_bench_c4.json 2> gpurun_out/${T}_bench_c4.err
python bench.py --workload c2 --cpu-frames 0 > gpurun_out/${T}_bench_c2.json 2> gpurun_out/${T}_bench_c2.err
python bench.py --workload c3 --cpu-frames 0 > gpurun_out/${T}_bench_c3.json 2> gpurun_out/${T}_bench_c3.err
python tools/show_bench.py gpurun_out/${T}_bench_*.json
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_stats|k_bilateral|k_upsharp|k_blend|k_scene" -s 20 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_${T}_a.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_bilateral|k_upsharp" -s 40 -c 4 -o gpurun_out/prof_${T} python bench.py --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_${T}_c.log 2>&1
du -sh gpurun_out
