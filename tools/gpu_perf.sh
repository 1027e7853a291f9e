set -x
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --kernel-breakdown > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_graph.json'))
print('TRAIN', d['ms_per_step'], d['value'], d['e2e']['value'], d['kernel_families'])
PY
