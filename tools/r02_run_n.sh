set -x
python -m pytest tests/test_gpu_stages.py tests/test_gpu_edges.py tests/test_gpu_guard.py -x -q 2>&1 | tail -2
AB_ARGS="--workload c1" bash tools/ab.sh r02n_c1 base
AB_ARGS="--workload c5" bash tools/ab.sh r02n_c5 base
