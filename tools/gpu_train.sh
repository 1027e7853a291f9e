set -x
timeout 600 python -m pytest tests/test_train_gpu.py -q -x -s > gpurun_out/train_gpu.log 2>&1
tail -6 gpurun_out/train_gpu.log
timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline --kernel-breakdown > gpurun_out/bench_train_graph.json 2> gpurun_out/bench_train_graph.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train_graph.json'))
print(d['ms_per_step'], d['value'], d['e2e'], d['roofline']['step'])
print(d['kernel_families'])
PY
grep -E "row_layernorm|transpose|attention|attn|pos_bias" gpurun_out/bench_train_graph.err
