# A/B timing of library variants on the C4 chain: bash tools/ab.sh tag variant1 variant2 ...
#   "base" = the product build, "env:NAME=VALUE" = the product build under that environment, else libvpb200_<variant>.so
T=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=video_processor_b200/libvpb200.so; envs=""
  case "$v" in base) ;; env:*) envs="${v#env:}"; v=$(echo "$v" | tr -c 'A-Za-z0-9_\n' '_') ;; *) lib=video_processor_b200/libvpb200_$v.so ;; esac
  env $envs VPB200_LIB=$PWD/$lib python bench.py --cpu-frames 0 --no-c5-latency --no-e2e ${AB_ARGS} > gpurun_out/${T}_$v.json 2> gpurun_out/${T}_$v.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${T}_$v.json"))
print("$v", round(d["value"],1), d.get("parity_check"), {k: v["avg_us"] for k,v in d["roofline"]["kernels"].items()})
PY
done
