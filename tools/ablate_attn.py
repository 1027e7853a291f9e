"""Ablation timing of the attention kernel (SGF_ATTN_DBG bit mask: 1 no exp, 2 no P store, 4 no bias add, 8 no max)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops
from tools.bench_ops import timeit

B, H, T = 8, 12, 936
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
bias = torch.randn(H, T, 960, device="cuda", generator=g).half()
out = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
for use_bias in (True, False):
    for dbg in (0, 1, 2, 4, 8, 3, 7, 15):
        os.environ["SGF_ATTN_DBG"] = str(dbg)
        fn = lambda: ops.attention(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, q_strides=(3 * D, T * 3 * D),  # noqa: E731
                                   k_strides=(3 * D, T * 3 * D), v_strides=(3 * D, T * 3 * D), o_strides=(D, T * D),
                                   bias=bias if use_bias else None)
        print(f"bias={use_bias} dbg={dbg:2d}  {timeit(fn, iters=10) * 1e3:.1f} us", flush=True)
