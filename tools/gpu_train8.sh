set -x
export NCCL_DEBUG=WARN
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_8gpu.json 2> gpurun_out/bench_train_8gpu.err
tail -c 400 gpurun_out/bench_train_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_train_8gpu.json').read().strip().splitlines()[-1])
print('TRAIN8', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['cuda_graph'])
PY
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer_8gpu.json 2> gpurun_out/bench_infer_8gpu.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_infer_8gpu.json').read().strip().splitlines()[-1])
print('INF8', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'])
PY
