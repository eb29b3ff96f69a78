set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
bash tools/ab.sh r02r base env:VPB200_NO_PRE_BGRX=1 base
AB_ARGS="--workload c2" bash tools/ab.sh r02r_c2 base
