set -x
AB_ARGS="--workload c5" bash tools/ab.sh r02n_c5 base th4_128
AB_ARGS="--workload c1" bash tools/ab.sh r02n_c1 base th2_64
VPB200_LIB=$PWD/video_processor_b200/libvpb200_th2_64.so python -m pytest tests/test_gpu_stages.py tests/test_gpu_edges.py tests/test_gpu_guard.py -x -q 2>&1 | tail -2
VPB200_LIB=$PWD/video_processor_b200/libvpb200_th4_128.so python -m pytest tests/test_gpu_stages.py tests/test_gpu_edges.py tests/test_gpu_guard.py -x -q 2>&1 | tail -2
