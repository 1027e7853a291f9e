set -x
timeout 900 python -m pytest tests/test_train_gpu.py -q -x > gpurun_out/row_ops.log 2>&1
tail -5 gpurun_out/row_ops.log
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
tail -3 gpurun_out/bench_train_graph.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_graph.json'))
print('TRAIN', d['ms_per_step'], d['value'], d['e2e']['value'], d['cuda_graph'])
PY
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_train_eager.json 2> gpurun_out/bench_train_eager.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_eager.json'))
print('TRAIN eager', d['ms_per_step'], d['value'], d['e2e']['value'], d['cuda_graph'])
PY
