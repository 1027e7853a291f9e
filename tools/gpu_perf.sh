set -x
(cd tools && timeout 300 python bench_attn.py) > gpurun_out/bench_attn.log 2>&1; grep -E "shape" gpurun_out/bench_attn.log | cut -c1-150
