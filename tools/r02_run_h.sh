set -x
python -m pytest tests/test_gpu_chain.py tests/test_gpu_headline.py tests/test_gpu_segments.py -x -q 2>&1 | tail -3
AB_ARGS="--workload c1" bash tools/ab.sh r02h_c1 base
AB_ARGS="--workload c5" bash tools/ab.sh r02h_c5 base
bash tools/ab.sh r02h_c4 base
