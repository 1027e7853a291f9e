set -x
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_8gpu.json 2> gpurun_out/bench_8gpu.err; echo "rc=$?"
tail -3 gpurun_out/bench_8gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_8gpu.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])
t=d.get('train_step')
print(t['n_gpus'], t['ms_per_step'], t['value'], t['cuda_graph'], t['grad_allreduce'])
PY
