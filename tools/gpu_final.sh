set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -3 gpurun_out/pytest_gpu.log; tail -2 gpurun_out/smoke.log; cut -c1-300 gpurun_out/bench_ref.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_final.json'))
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'])
print(d['roofline']['frac'], d['roofline'].get('attention'), d['roofline'].get('step'))
print(d.get('parity')); print(d.get('gpu_eager_context')); print(d['cpu_baseline']['value'], d['clocks'])
print(d['train_step']['ms_per_step'], d['train_step']['value'])
PY
