set -x
timeout 300 python -m pytest tests/test_ops_gpu.py tests/test_train_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
timeout 200 python bench.py --config 3 --steps 8 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('TRAIN', d['ms_per_step'], d['value'])"
