# the command list behind profiles/r01g_* (run on the GPU box: gpurun -- 'bash tools/record_round.sh')
set -x
T=r01g
for w in c4 c1 c1w c2 c3 c5; do
  extra="--cpu-frames 0"; [ $w = c4 ] && extra=""; [ $w = c1w ] && extra="--cpu-frames 24"
  python bench.py --workload $w $extra > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
python bench.py --bilateral-mode selective --frames 64 --cpu-frames 6 > gpurun_out/${T}_bench_c4_selective.json 2> gpurun_out/${T}_sel.err
python bench.py --bilateral-mode multiscale --frames 64 --cpu-frames 6 > gpurun_out/${T}_bench_c4_multiscale.json 2> gpurun_out/${T}_ms.err
python bench.py --adaptive-sigma --frames 64 --cpu-frames 6 > gpurun_out/${T}_bench_c4_adaptive_sigma.json 2> gpurun_out/${T}_as.err
python bench.py --flow --frames 32 --cpu-frames 3 > gpurun_out/${T}_bench_c4_flow.json 2> gpurun_out/${T}_flow.err
python bench.py --bilateral-mode multiscale --adaptive-sigma --flow --frames 32 --cpu-frames 3 > gpurun_out/${T}_bench_c4_reference_defaults.json 2> gpurun_out/${T}_rd.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
# launch lists (share of the step per kernel; only after the plain runs above exited 0)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_stats|k_bilateral|k_upsharp|k_blend|k_scene" -s 20 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e > gpurun_out/ncu_${T}_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv --log-file gpurun_out/${T}_launches_reference_defaults.csv python bench.py --steps 1 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e --bilateral-mode multiscale --adaptive-sigma --flow > gpurun_out/ncu_${T}_b.log 2>&1
python tools/show_bench.py gpurun_out/${T}_bench_*.json
