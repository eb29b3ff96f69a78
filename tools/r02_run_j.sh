set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/ab.sh r02j base env:VPB200_NO_BGRX=1 minb3 base
AB_ARGS="--workload c2" bash tools/ab.sh r02j_c2 base
