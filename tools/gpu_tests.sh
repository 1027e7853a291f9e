set -x
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline --no-train-record > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; cut -c1-330 gpurun_out/bench_a.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_a.json'))
print(d['kernel_families'])
PY
