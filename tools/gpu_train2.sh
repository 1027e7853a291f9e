set -x
export NCCL_DEBUG=WARN
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --config 3 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_train_2gpu.json 2> gpurun_out/bench_train_2gpu.err
tail -c 1500 gpurun_out/bench_train_2gpu.json; tail -5 gpurun_out/bench_train_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --config 3 --steps 10 --warmup 3 --no-cpu-baseline --no-graph > gpurun_out/bench_train_2gpu_eager.json 2> gpurun_out/bench_train_2gpu_eager.err
tail -c 600 gpurun_out/bench_train_2gpu_eager.json; tail -3 gpurun_out/bench_train_2gpu_eager.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_infer_2gpu.json 2> gpurun_out/bench_infer_2gpu.err
tail -c 700 gpurun_out/bench_infer_2gpu.json
