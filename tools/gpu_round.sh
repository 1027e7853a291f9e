set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python bench.py --config 3 --steps 10 --warmup 3 > gpurun_out/bench_train_final.json 2> gpurun_out/bench_train_final.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cat gpurun_out/bench_ref.json; cat gpurun_out/bench_final.json; cat gpurun_out/bench_train_final.json | cut -c1-700
timeout 600 python bench.py --config 4 --no-cpu-baseline > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 900 python bench.py --config 5 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_cfg5.json 2> gpurun_out/bench_cfg5.err
python - <<'PY'
import json
for n in ("cfg4", "cfg5"):
    try:
        d = json.load(open(f"gpurun_out/bench_{n}.json"))
        print(n, d["ms_per_step"], d["value"], d["roofline"]["step"])
    except Exception as e:
        print(n, "failed", e)
PY
