set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --kernel-breakdown > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
grep -E "row_|gelu_ln|colsum|upsample|attn|attention" gpurun_out/bench_train_graph.err | head -30
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_graph.json'))
print('TRAIN', d['ms_per_step'], d['value'], d['e2e']['value'], d['cuda_graph'])
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_infer.json 2> gpurun_out/bench_infer.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_infer.json'))
print('INFER', d['ms_per_step'], d['value'], d['e2e']['value'])
PY
