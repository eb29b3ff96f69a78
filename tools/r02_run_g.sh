set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/ab.sh r02g base i2f1 base
AB_ARGS="--workload c2" bash tools/ab.sh r02g_c2 base
AB_ARGS="--workload c1" bash tools/ab.sh r02g_c1 base
AB_ARGS="--workload c5" bash tools/ab.sh r02g_c5 base
