timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_criterion_gpu.py tests/test_train_ops_gpu.py -q -x -k "upsample or criterion or ce" > gpurun_out/seg_tests.log 2>&1; echo "rc=$?" >> gpurun_out/seg_tests.log
tail -5 gpurun_out/seg_tests.log
python bench.py --config 4 --no-cpu-baseline --no-train-record 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg4', d['ms_per_step'], d['kernel_families'].get('upsample_argmax'))"
python bench.py --config 3 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('cfg3', d['ms_per_step'], {k:v for k,v in d['kernel_families'].items() if 'upsample' in k})"
