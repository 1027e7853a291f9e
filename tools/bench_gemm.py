"""GEMM variant sweep (tile width / pipeline depth) on the transformer and stem shapes."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from ifseg_b200 import ops, _lib
from tools.bench_ops import timeit

def main():
    lib = _lib.load()
    shapes = [(7488, 2304, 768), (7488, 768, 768), (7488, 3072, 768), (7488, 768, 3072), (7200, 1024, 256),
              (7488, 9216, 768), (8192, 8192, 8192), (300, 512, 256), (7200, 1024, 256)]
    variants = [(128, 3), (1256, 1), (1256, 0), (0, 0)]
    for (M, N, K) in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ref = None
        row = {}
        for bn, st in variants:
            lib.sgf_gemm_force_variant(bn, st)
            ms = timeit(lambda: ops.gemm(a, b, out), iters=10)
            if ref is None:
                ref = out.float().clone()
            else:
                err = (out.float() - ref).abs().max().item()
                assert err == 0, (bn, st, err)
            row[f"bn{bn}_s{st}"] = round(2 * M * N * K / ms / 1e9)
        lib.sgf_gemm_force_variant(0, 0)
        ms_t = timeit(lambda: torch.matmul(a, b.t(), out=out), iters=10)
        row["cublas"] = round(2 * M * N * K / ms_t / 1e9)
        print((M, N, K), row, flush=True)

if __name__ == "__main__":
    main()
