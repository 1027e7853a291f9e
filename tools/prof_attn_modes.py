"""One backward launch per bias mode (none / bias / bias+dbias) at the cfg-3 encoder shape, for an ncu launch list."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops

B, H, Tq, Tk = 8, 12, 1115, 1115
D = H * 64
q = torch.randn(B, Tq, D, device="cuda").bfloat16() * 0.3
k = torch.randn(B, Tk, D, device="cuda").bfloat16()
v = torch.randn(B, Tk, D, device="cuda").bfloat16()
do = torch.randn(B, Tq, D, device="cuda").bfloat16() * 0.1
Tkp = (Tk + 63) // 64 * 64
bias = torch.randn(H, Tq, Tkp, device="cuda").half()
out = torch.empty(B, Tq, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, Tq, device="cuda")
delta = torch.empty(B, H, Tq, device="cuda")
dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
dbias = torch.zeros(H, Tq, Tkp, device="cuda")
bias_t = ops.transpose_bias(bias, Tk)
st = dict(B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D), v_strides=(D, Tk * D), o_strides=(D, Tq * D))
for rep in range(3):
    if rep == 2:
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    for bb, db in ((None, None), (bias, None), (bias, dbias)):
        ops.attention(q, k, v, out, bias=bb, lse=lse, **st)
        ops.attention_bwd(q, k, v, out, do, dq, dk, dv, do_strides=(D, Tq * D), dq_strides=(D, Tq * D),
                          dk_strides=(D, Tk * D), dv_strides=(D, Tk * D), lse=lse, delta=delta, bias=bb, dbias=db,
                          bias_t=bias_t if bb is not None else None, **st)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
