set -x
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_new.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'])
print(d['roofline']['frac'], d['roofline']['attention'], d['roofline']['step'])
print(d['parity']); print(d['gpu_eager_context']); print(d['cpu_baseline']['value'])
print(d['train_step']['ms_per_step'], d['train_step']['value'])
PY
