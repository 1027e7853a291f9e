"""globaltimer timeline of the first and the last cluster of a gemm_ts launch (debug hook sgf_debug_set_gemm_trace)."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import _lib, ops

os.environ["SGF_GEMM_FAMILY"] = "ts"
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(0)
ROLES = ["cta (start, prologue done, pdl_wait done, wg0 last store read, exit)", "producer A: k-block issue",
         "mma: k-block operands landed", "mma: tile committed", "epilogue wg0: box store issued", "-",
         "epilogue: tmem_full seen (tile*2+wg)", "-"]
for (M, N, K, f32, bn) in [(7488, 768, 64, False, 256), (7488, 768, 768, True, 384), (7488, 768, 768, True, 256), (7200, 256, 1024, False, 128)]:
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = torch.empty(M, N, device="cuda", dtype=torch.float32 if f32 else torch.bfloat16)
    os.environ["SGF_GEMM_TS_BN"] = str(bn)
    for _ in range(3):
        ops.gemm(a, b, out, bias=bias)
    torch.cuda.synchronize()
    tr = torch.zeros(2 * 8 * 64, dtype=torch.int64, device="cuda")
    lib.sgf_debug_set_gemm_trace(tr.data_ptr())
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    ops.gemm(a, b, out, bias=bias)
    e.record()
    torch.cuda.synchronize()
    lib.sgf_debug_set_gemm_trace(None)
    t = tr.view(2, 8, 64).cpu()
    t0 = int(t[t > 0].min())
    print(f"==== M={M} N={N} K={K} {'f32' if f32 else 'bf16'} BN={bn}   event-timed {s.elapsed_time(e) * 1e3:.1f} us; times in ns from the first CTA start")
    for slot in range(2):
        print(f"-- {'first' if slot == 0 else 'last'} cluster")
        for r in range(8):
            vals = [int(v) - t0 for v in t[slot, r] if v > 0]
            print(f"   {ROLES[r]:62s} " + " ".join(str(v) for v in vals[:26]))
