set -x
timeout 300 python tools/trace_attn.py > gpurun_out/trace_attn.log 2>&1
tail -5 gpurun_out/trace_attn.log
