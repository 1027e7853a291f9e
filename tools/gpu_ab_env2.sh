# usage: bash tools/gpu_ab_env2.sh VAR v1 v2 ...   -> inference bench with VAR=each value
VAR=$1; shift
for v in "$@"; do
  env ${VAR}=${v} python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-train-record > gpurun_out/ab_${v}.json 2> gpurun_out/ab.err
  python - <<PY
import json
d=json.load(open("gpurun_out/ab_${v}.json"))
print("${VAR}=${v}", round(d["ms_per_step"],4), round(d["value"],1), d["roofline"].get("row_layernorm",{}).get("ms_in_step"))
PY
done
