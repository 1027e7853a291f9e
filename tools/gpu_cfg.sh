set -x
timeout 600 python bench.py --config 4 --no-cpu-baseline --no-train-record > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 900 python bench.py --config 5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
timeout 900 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python - <<'PY'
import json
for n in ("cfg4", "cfg5", "cfg3", "ref"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d["ms_per_step"], d["value"], d.get("roofline", {}).get("step"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(n, "failed", e)
PY
