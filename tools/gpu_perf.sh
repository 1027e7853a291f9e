set -x
SGF_GEMM_PAIR=2 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --kernel-breakdown > gpurun_out/bench_inf2.json 2> gpurun_out/bench_inf2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_inf2.json'))
print('INF mode2', d['ms_per_step'], d['value'], d['kernel_families']['gemm_tcgen05'])
PY
grep gemm gpurun_out/bench_inf2.err | head -12
SGF_GEMM_PAIR=2 timeout 600 python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train2.json 2> gpurun_out/bench_train2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_train2.json'))
print('TRAIN mode2', d['ms_per_step'], d['value'], d['kernel_families']['gemm_tcgen05'])
PY
