set -x
timeout 900 python -m pytest tests/test_train_ops_gpu.py tests/test_train_gpu.py -q -x > gpurun_out/row_ops.log 2>&1
tail -8 gpurun_out/row_ops.log
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --kernel-breakdown > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_graph.json'))
print('TRAIN', d['ms_per_step'], d['value'], d['e2e']['value'])
PY
grep row_layernorm gpurun_out/bench_train_graph.err
