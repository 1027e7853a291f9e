set -x
timeout 600 python -m pytest tests/test_train_ops_gpu.py tests/test_train_gpu.py -x -q > gpurun_out/quick.log 2>&1; tail -6 gpurun_out/quick.log
timeout 300 python tools/bench_attn.py 2>&1 | grep shape
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_graph.json'))
print('TRAIN', d['ms_per_step'], d['value'], d['e2e']['value'], d['cuda_graph'])
PY
