set -x
python -m pytest tests/test_gpu_stages.py tests/test_gpu_chain.py -x -q 2>&1 | tail -2
VPB200_LIB=$PWD/video_processor_b200/libvpb200_cls5.so python -m pytest tests/test_gpu_stages.py tests/test_gpu_chain.py tests/test_gpu_headline.py -x -q 2>&1 | tail -2
bash tools/ab.sh r02l base cls5 cls5b3 base
