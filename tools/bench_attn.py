"""Attention micro-benchmark: with / without the additive bias stream (forward and backward)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops
from bench_ops import timeit

for (B, H, Tq, Tk, causal) in [(8, 12, 936, 936, False), (8, 12, 1115, 1115, False), (8, 12, 901, 901, True)]:
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
    st = dict(B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D), v_strides=(D, Tk * D), o_strides=(D, Tq * D))
    bias32 = bias.float()
    dbias = torch.zeros(H, Tq, Tkp, device="cuda")
    bias_t = ops.transpose_bias(bias, Tk)
    for name, bb in (("bias+dbias", bias), ("bias", bias), ("nobias", None)):
        ms = timeit(lambda: ops.attention(q, k, v, out, bias=bb, causal=causal, lse=lse, **st))
        ms_b = timeit(lambda: ops.attention_bwd(q, k, v, out, do, dq, dk, dv, do_strides=(D, Tq * D), dq_strides=(D, Tq * D),
                                                dk_strides=(D, Tk * D), dv_strides=(D, Tk * D), lse=lse, delta=delta, bias=bb,
                                                causal=causal, dbias=dbias if name == "bias+dbias" else None,
                                                bias_t=bias_t if bb is not None else None, **st))
        pairs = Tq * Tk if not causal else Tq * (Tq + 1) / 2
        print(dict(shape=(B, H, Tq, Tk, causal), mode=name, fwd_ms=round(ms, 4), fwd_tflops=round(4 * B * H * pairs * 64 / ms / 1e9, 1),
                   bwd_ms=round(ms_b, 4), bwd_tflops=round(10 * B * H * pairs * 64 / ms_b / 1e9, 1)), flush=True)
