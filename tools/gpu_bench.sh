set -x
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_new.json 2> gpurun_out/bench_new.err; echo "rc=$?"
tail -5 gpurun_out/bench_new.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_new.json'))
print(d['ms_per_step'], d['value'], d['e2e'])
print(json.dumps(d['roofline'])[:900])
print(d.get('parity'))
print(json.dumps(d.get('train_step'))[:1200])
print(d.get('cpu_baseline'))
PY
