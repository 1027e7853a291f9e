"""Per-role clock64 timeline of a few attention CTAs (debug hook sgf_debug_set_attention_trace)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import _lib, ops

B, H, T = 8, 12, 936
D = H * 64
g = torch.Generator(device="cuda").manual_seed(0)
qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
bias = torch.zeros(H, T, 960, device="cuda")
bias[:, :, :T] = torch.randn(H, T, T, device="cuda", generator=g)
bias = bias.half()
out = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
lib = _lib.load()


def run(use_bias):
    ops.attention(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, q_strides=(3 * D, T * 3 * D),
                  k_strides=(3 * D, T * 3 * D), v_strides=(3 * D, T * 3 * D), o_strides=(D, T * D),
                  bias=bias if use_bias else None)


for use_bias in (True, False):
    run(use_bias)
    torch.cuda.synchronize()
    tr = torch.zeros(4 * 4 * 32 * 4, dtype=torch.int64, device="cuda")
    lib.sgf_debug_set_attention_trace(tr.data_ptr())
    run(use_bias)
    torch.cuda.synchronize()
    lib.sgf_debug_set_attention_trace(None)
    t = tr.view(4, 4, 32, 4).cpu()
    print("==== bias" if use_bias else "==== no bias")
    for cta in range(4):
        t0 = int(t[cta][t[cta] > 0].min()) if (t[cta] > 0).any() else 0
        print(f"-- cta slot {cta}")
        names = ["S-issuer  (loop top, k_full ok, s_empty ok, issued)", "PV-issuer (loop top, v_full ok, p_full ok, issued)",
                 "softmax w0 (top, s_full ok, max0 done, exp0 done)", "softmax w3 (top, s_full ok, max0 done, exp0 done)"]
        for role in range(4):
            print("  " + names[role])
            for j in range(15):
                print("    j=%2d " % j + " ".join("%7d" % (int(v) - t0 if v > 0 else -1) for v in t[cta, role, j]))
