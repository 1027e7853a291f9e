"""Micro-benchmarks of the individual kernels (CUDA events, L2-cold rotation)."""
import json
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    res = []
    for (M, N, K) in [(7488, 2304, 768), (7488, 768, 768), (7488, 3072, 768), (7488, 768, 3072), (7200, 768, 1024),
                      (7200, 1024, 256), (7200, 256, 1024), (115200, 64, 64), (115200, 256, 64), (28800, 512, 128),
                      (8192, 8192, 8192)]:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        ms = timeit(lambda: ops.gemm(a, b, out))
        ms_t = timeit(lambda: torch.matmul(a, b.t(), out=out))
        res.append(dict(op="gemm", M=M, N=N, K=K, ms=ms, tflops=2 * M * N * K / ms / 1e9, cublas_ms=ms_t,
                        cublas_tflops=2 * M * N * K / ms_t / 1e9))
        print(res[-1], flush=True)
    for (n, h, w, c) in [(8, 30, 30, 256), (8, 60, 60, 128), (8, 120, 120, 64)]:
        x = torch.randn(n, h, w, c, device="cuda").bfloat16()
        wt = torch.randn(c, 3, 3, c, device="cuda").bfloat16()
        sc, bi = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
        ms = timeit(lambda: ops.conv3x3_s1(x, wt, sc, bi))
        res.append(dict(op="conv3x3", n=n, h=h, w=w, c=c, ms=ms, tflops=2 * n * h * w * c * c * 9 / ms / 1e9))
        print(res[-1], flush=True)
    for (B, H, Tq, Tk, causal) in [(8, 12, 936, 936, False), (8, 12, 901, 901, True), (8, 12, 901, 936, False)]:
        D = H * 64
        q = torch.randn(B, Tq, D, device="cuda").bfloat16() * 0.3
        k = torch.randn(B, Tk, D, device="cuda").bfloat16()
        v = torch.randn(B, Tk, D, device="cuda").bfloat16()
        Tkp = (Tk + 63) // 64 * 64
        bias = torch.randn(H, Tq, Tkp, device="cuda").half()
        out = torch.empty(B, Tq, D, device="cuda", dtype=torch.bfloat16)
        f = lambda: ops.attention(q, k, v, out, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D),
                                  v_strides=(D, Tk * D), o_strides=(D, Tq * D), bias=bias, causal=causal)
        ms = timeit(f)
        fl = 4 * B * H * Tq * Tk * 64 * (0.5 if causal else 1.0)
        res.append(dict(op="attention", B=B, H=H, Tq=Tq, Tk=Tk, causal=causal, ms=ms, tflops=fl / ms / 1e9))
        print(res[-1], flush=True)
    for (rows, D) in [(7488, 768), (7488, 3072)]:
        x = torch.randn(rows, D, device="cuda").bfloat16()
        g, b = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
        o = torch.empty_like(x)
        ms = timeit(lambda: ops.row_layernorm(x, ln2=(g, b), out2=o))
        res.append(dict(op="layernorm", rows=rows, D=D, ms=ms, gbs=rows * D * 4 / ms / 1e6))
        print(res[-1], flush=True)
    logits = torch.randn(8, 901, 15, device="cuda")
    ms = timeit(lambda: ops.upsample_argmax(logits, 30, 30, 480, 480))
    res.append(dict(op="upsample_argmax", ms=ms, gbs=8 * 480 * 480 * 8 / ms / 1e6))
    print(res[-1], flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_ops.json", "w"), indent=1)


if __name__ == "__main__":
    main()
