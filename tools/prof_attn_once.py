"""One cfg-2 encoder-shape attention launch (after a warm-up) for ncu --set full."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops

B, H, T = 8, 12, 936
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
bias = torch.zeros(H, T, 960, device="cuda")
bias[:, :, :T] = torch.randn(H, T, T, device="cuda", generator=g)
bias = bias.half()
out = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
for use_bias in (True, False, True, False):
    ops.attention(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, q_strides=(3 * D, T * 3 * D),
                  k_strides=(3 * D, T * 3 * D), v_strides=(3 * D, T * 3 * D), o_strides=(D, T * D),
                  bias=bias if use_bias else None)
torch.cuda.synchronize()
