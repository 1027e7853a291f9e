set -x
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; echo "rc=$?"
tail -5 gpurun_out/bench_2gpu.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_2gpu.json'))
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e'])
t=d.get('train_step')
print(t['n_gpus'], t['ms_per_step'], t['value'], t['cuda_graph'], t['grad_allreduce'])
PY
