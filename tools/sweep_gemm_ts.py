"""Cost structure of the TMA-store pair GEMM: time vs K (slope = per-k-block cost, intercept = fixed overhead) per tile
width and output type, L2-cold and L2-warm."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops
from tools.bench_ops import timeit


def timeit_warm(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


os.environ["SGF_GEMM_FAMILY"] = "ts"
g = torch.Generator(device="cuda").manual_seed(0)
for (M, N) in [(7488, 768), (7488, 2304), (7488, 3072)]:
    for f32 in (False, True):
        for bn in (128, 192, 256):
            if N % bn:
                continue
            os.environ["SGF_GEMM_TS_BN"] = str(bn)
            row = []
            for K in (64, 256, 768, 1536, 3072):
                a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
                b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
                bias = torch.randn(N, device="cuda", generator=g)
                out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
                fn = lambda: ops.gemm(a, b, out, bias=bias)  # noqa: E731
                row.append(f"K={K}: {timeit(fn, iters=10) * 1e3:.1f}/{timeit_warm(fn) * 1e3:.1f}")
            print(f"M={M} N={N} {'f32' if f32 else 'bf16'} BN={bn}  cold/warm us  " + "  ".join(row), flush=True)
