# round-2, one 8-GPU box: topology, copy ceiling at N=1,2,4,8 (plain and NUMA-bound), the bench line at N=2,4,8
set -x
mkdir -p gpurun_out
bash tools/box_topology.sh > gpurun_out/r02_topo_n8box.txt 2>&1
python tools/copy_peak.py --gpus 1,2,4,8 --modes d2h,both --seconds 0.8 --out gpurun_out/r02_copy_peak_n8box.json > gpurun_out/r02_copy_peak_n8box.log 2>&1
python tools/copy_peak.py --gpus 2,8 --modes both --seconds 0.8 --bind --out gpurun_out/r02_copy_peak_n8box_bind.json > gpurun_out/r02_copy_peak_n8box_bind.log 2>&1
for n in 8 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --steps 3 --warmup 3 --cpu-frames 0 --no-c5-latency > gpurun_out/r02c_bench_n$n.json 2> gpurun_out/r02c_bench_n$n.err
done
python tools/show_bench.py gpurun_out/r02c_bench_n*.json
cat gpurun_out/r02_copy_peak_n8box.log gpurun_out/r02_copy_peak_n8box_bind.log
