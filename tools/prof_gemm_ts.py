"""A few launches of the TMA-store pair GEMM for ncu (one warm-up launch each, then the profiled one)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops

os.environ["SGF_GEMM_FAMILY"] = "ts"
g = torch.Generator(device="cuda").manual_seed(0)
cases = []
for (M, N, K, f32, bn) in [(7488, 768, 64, False, 256), (7488, 768, 768, True, 256), (7488, 2304, 768, False, 256),
                           (7488, 3072, 768, False, 256), (7488, 768, 3072, True, 256), (7488, 768, 768, True, 192),
                           (115200, 256, 64, False, 256)]:
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    cases.append((bn, a, b, bias, out))
for rep in range(2):
    for bn, a, b, bias, out in cases:
        os.environ["SGF_GEMM_TS_BN"] = str(bn)
        ops.gemm(a, b, out, bias=bias)
    torch.cuda.synchronize()
