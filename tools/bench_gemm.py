"""A/B of the GEMM kernel families (SGF_GEMM_FAMILY = tile | ts) on the call sites of the cfg-2 forward, with
their real epilogues, checked against each other bit for bit and timed L2-cold next to cuBLAS (plain GEMM only)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops
from tools.bench_ops import timeit


def run(name, fn, out, flops, fams=("tile", "ts")):
    row, ref = {}, None
    for fam in fams:
        os.environ["SGF_GEMM_FAMILY"] = fam
        try:
            ms = timeit(fn, iters=10)
        except Exception as e:  # noqa: BLE001
            row[fam] = f"err {str(e)[:40]}"
            continue
        cur = out.float().clone()
        if ref is None:
            ref = cur
        else:
            d = (cur - ref).abs().max().item()
            row[fam + "_maxdiff"] = d
        row[fam] = f"{ms * 1e3:.1f}us {flops / ms / 1e9:.0f}TF"
    os.environ.pop("SGF_GEMM_FAMILY", None)
    ms = timeit(fn, iters=10)
    row["auto"] = f"{ms * 1e3:.1f}us {flops / ms / 1e9:.0f}TF"
    print(name, row, flush=True)


def main():
    g = torch.Generator(device="cuda").manual_seed(0)
    rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
    for (tag, M, N, K) in [("out_proj", 7488, 768, 768), ("cross_q", 7208, 768, 768), ("fc2", 7488, 768, 3072),
                           ("qkv", 7488, 2304, 768), ("fc1", 7488, 3072, 768), ("image_proj", 7200, 768, 1024)]:
        a, b = rn(M, K).bfloat16(), (rn(N, K) * 0.05).bfloat16()
        bias = rn(N)
        if tag == "fc2":
            x = rn(M, N)
            stats = torch.empty(M, K // 64, 2, device="cuda")
            f = a.float().view(M, K // 64, 64)
            stats[..., 0], stats[..., 1] = f.sum(-1), (f * f).sum(-1)
            u = rn(N)
            out = x  # in place, as the engine calls it (x += fc2(...)); values drift over the timing loop, which is harmless
            run(tag, lambda: ops.gemm(a, b, out, bias=bias, residual=out, rownorm=(stats, u, K)), out, 2.0 * M * N * K)
        elif tag in ("out_proj", "image_proj"):
            out = torch.empty(M, N, device="cuda")
            run(tag, lambda: ops.gemm(a, b, out, bias=bias), out, 2.0 * M * N * K)
        elif tag == "fc1":
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            run(tag, lambda: ops.gemm(a, b, out, bias=bias, act=ops.ACT_GELU), out, 2.0 * M * N * K)
        else:
            out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            run(tag, lambda: ops.gemm(a, b, out, bias=bias, alpha=0.088, alpha_cols=768), out, 2.0 * M * N * K)
        o2 = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: torch.matmul(a, b.t(), out=o2), iters=10)
        print(f"   cuBLAS plain {ms * 1e3:.1f}us {2.0 * M * N * K / ms / 1e9:.0f}TF", flush=True)
    # stem: 1x1 convolutions as GEMMs with the BN/ReLU(/residual) epilogue, 3x3 implicit GEMM
    for (tag, M, N, K, res) in [("l3.conv1", 7200, 256, 1024, False), ("l3.conv3", 7200, 1024, 256, True),
                                ("l2.conv1", 28800, 128, 512, False), ("l2.conv3", 28800, 512, 128, True),
                                ("l1.conv1", 115200, 64, 256, False), ("l1.conv3", 115200, 256, 64, True),
                                ("conv1_7x7", 460800, 64, 168, False)]:
        a, b = rn(M, K).bfloat16(), (rn(N, K) * 0.05).bfloat16()
        sc, bi = rn(N), rn(N)
        idn = rn(M, N).bfloat16() if res else None
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        run(tag, lambda: ops.gemm(a, b, out, scale=sc, bias=bi, act=ops.ACT_RELU, residual=idn), out, 2.0 * M * N * K,
            fams=("tile", "ts"))
    for (tag, n, h, c) in [("l3.conv2", 8, 30, 256), ("l2.conv2", 8, 60, 128), ("l1.conv2", 8, 120, 64)]:
        x = rn(n, h, h, c).bfloat16()
        w = (rn(c, 9 * c) * 0.05).bfloat16()
        sc, bi = rn(c), rn(c)
        out = torch.empty(n, h, h, c, device="cuda", dtype=torch.bfloat16)
        run(tag, lambda: ops.conv3x3_s1(x, w, sc, bi, out=out), out, 2.0 * n * h * h * c * 9 * c, fams=("tile", "ts"))


if __name__ == "__main__":
    main()
