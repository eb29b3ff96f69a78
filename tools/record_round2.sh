# the command list behind profiles/r02z_* (run on one B200: gpurun -- 'bash tools/record_round2.sh')
set -x
T=${1:-r02z}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${T}_gputest.log 2>&1; echo rc=$? >> gpurun_out/${T}_gputest.log
tail -2 gpurun_out/${T}_gputest.log
for w in c4 c1 c1w c2 c3 c5; do
  extra="--cpu-frames 0"; [ $w = c4 ] && extra=""; [ $w = c1w ] && extra="--cpu-frames 24"
  python bench.py --workload $w $extra > gpurun_out/${T}_bench_$w.json 2> gpurun_out/${T}_bench_$w.err
done
python bench.py --bilateral-mode selective --frames 64 --cpu-frames 0 > gpurun_out/${T}_bench_c4_selective.json 2> gpurun_out/${T}_sel.err
python bench.py --bilateral-mode multiscale --frames 64 --cpu-frames 0 > gpurun_out/${T}_bench_c4_multiscale.json 2> gpurun_out/${T}_ms.err
python bench.py --adaptive-sigma --frames 64 --cpu-frames 0 > gpurun_out/${T}_bench_c4_adaptive_sigma.json 2> gpurun_out/${T}_as.err
python bench.py --flow --frames 32 --cpu-frames 0 > gpurun_out/${T}_bench_c4_flow.json 2> gpurun_out/${T}_flow.err
python bench.py --bilateral-mode multiscale --adaptive-sigma --flow --frames 32 --cpu-frames 0 > gpurun_out/${T}_bench_c4_reference_defaults.json 2> gpurun_out/${T}_rd.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference_arm.json 2> gpurun_out/${T}_ref.err
python tools/show_bench.py gpurun_out/${T}_bench_*.json
# launch lists (share of the step per kernel; only after the plain runs above exited 0)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_stats|k_bilateral|k_upsharp|k_blend|k_scene" -s 20 -c 200 --csv --log-file gpurun_out/${T}_launches.csv python bench.py --steps 2 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_${T}_a.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 4000 -c 600 --csv --log-file gpurun_out/${T}_launches_reference_defaults.csv python bench.py --steps 1 --warmup 3 --frames 16 --cpu-frames 0 --no-e2e --no-c5-latency --bilateral-mode multiscale --adaptive-sigma --flow > gpurun_out/ncu_${T}_b.log 2>&1
# full captures: the headline kernels of one C4 frame, the integer-ratio bicubic at C1 and C5
ncu --set full --clock-control none --import-source on -k regex:"k_stats|k_bilateral|k_upsharp" -s 40 -c 4 -o gpurun_out/prof_${T} python bench.py --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame > gpurun_out/ncu_${T}_c.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bicubic" -s 20 -c 1 -o gpurun_out/prof_${T}_c1 python bench.py --workload c1 --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_${T}_c1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_bicubic" -s 20 -c 1 -o gpurun_out/prof_${T}_c5 python bench.py --workload c5 --steps 1 --warmup 3 --frames 8 --cpu-frames 0 --no-e2e --per-frame > gpurun_out/ncu_${T}_c5.log 2>&1
# the widened kernels whose committed rows predate their last rewrite (k_ms_blend, k_pyr_down, k_scene_*, detail / varsharp / flow)
ncu --set full --clock-control none -k regex:"k_ms_blend|k_pyr_down|k_scene|k_detail|k_texture|k_sigma_map|k_varsharp|k_joint|k_update_matrices|k_box_solve|k_polyexp|k_warp|k_reliability|k_blend_masked" --launch-skip 300 -c 40 -o /tmp/prof_${T}_widened python bench.py --steps 1 --warmup 2 --frames 8 --cpu-frames 0 --no-e2e --no-c5-latency --per-frame --bilateral-mode multiscale --adaptive-sigma --flow --scene-detect > gpurun_out/ncu_${T}_d.log 2>&1
# gpurun brings back at most 64 MiB: the widened capture goes home as its raw-page CSV only
ncu -i /tmp/prof_${T}_widened.ncu-rep --page raw --csv > gpurun_out/${T}_widened_raw.csv 2> /dev/null
ls -la gpurun_out/*.ncu-rep /tmp/*.ncu-rep; du -sh gpurun_out
