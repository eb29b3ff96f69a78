set -x
python -m pytest tests/test_gpu_chain.py tests/test_gpu_headline.py tests/test_gpu_segments.py -x -q 2>&1 | tail -2
bash tools/ab.sh r02q base env:VPB200_NO_STEADY=1 base
