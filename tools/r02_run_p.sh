set -x
python -m pytest tests/test_gpu_stages.py tests/test_gpu_chain.py tests/test_gpu_headline.py tests/test_gpu_edges.py -x -q 2>&1 | tail -2
bash tools/ab.sh r02p base p5int base
