set -x
SGF_ATTN_DBG=15 timeout 300 python tools/trace_attn.py > gpurun_out/trace_attn_dbg15.log 2>&1
