"""Attention forward micro-benchmark at the BASELINE shapes (CUDA events, L2 flushed between launches)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops
from bench_ops import timeit

tag = "attention_tcgen05"
NSHAPES = int(os.environ.get("BENCH_SHAPES", "5"))
for (B, H, Tq, Tk, causal) in [(8, 12, 936, 936, False), (8, 12, 1115, 1115, False), (8, 12, 901, 901, True), (8, 12, 901, 936, False),
                               (4, 16, 1815, 1815, False)][:NSHAPES]:
    D = H * 64
    q = torch.randn(B, Tq, D, device="cuda").bfloat16() * 0.3
    k = torch.randn(B, Tk, D, device="cuda").bfloat16()
    v = torch.randn(B, Tk, D, device="cuda").bfloat16()
    Tkp = (Tk + 63) // 64 * 64
    bias = torch.randn(H, Tq, Tkp, device="cuda").half()
    out = torch.empty(B, Tq, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, Tq, device="cuda")
    st = dict(B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D), v_strides=(D, Tk * D), o_strides=(D, Tq * D))
    for name, bb in (("bias", bias), ("nobias", None)):
        ms = timeit(lambda: ops.attention(q, k, v, out, bias=bb, causal=causal, lse=lse, **st))
        pairs = Tq * Tk if not causal else Tq * (Tq + 1) / 2
        print(dict(kernel=tag, shape=(B, H, Tq, Tk, causal), mode=name, fwd_us=round(ms * 1e3, 1),
                   fwd_tflops=round(4 * B * H * pairs * 64 / ms / 1e9, 1)), flush=True)
