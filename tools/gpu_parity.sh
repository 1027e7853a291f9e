set -x
rm -f gpurun_out/upsample_arith.jsonl
timeout 600 python -m pytest tests/test_ops_gpu.py -q -s -k "upsample_argmax" > gpurun_out/upsample_gpu.log 2>&1; tail -3 gpurun_out/upsample_gpu.log | cut -c1-300
python - <<'PY'
import json
for l in open('gpurun_out/upsample_arith.jsonl'):
    d=json.loads(l); print(d['shape'], {k:(v['vs_aten_cuda'],v['vs_aten_cuda_contiguous']) for k,v in d['mismatches'].items()})
PY
timeout 900 python -m pytest tests/test_criterion_gpu.py tests/test_ddp_gpu.py -q -x > gpurun_out/crit_gpu.log 2>&1; tail -30 gpurun_out/crit_gpu.log | cut -c1-300
timeout 900 python -m pytest tests/test_train_gpu.py tests/test_model_gpu.py tests/test_train_ops_gpu.py -q -x > gpurun_out/train_gpu.log 2>&1; tail -8 gpurun_out/train_gpu.log | cut -c1-300
