"""One launch of every GEMM call-site shape per kernel family (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from ifseg_b200 import ops

g = torch.Generator(device="cuda").manual_seed(0)
rn = lambda *s: torch.randn(*s, device="cuda", generator=g)  # noqa: E731
fams = sys.argv[1:] or ["tile", "persist"]
cases = []
for (tag, M, N, K) in [("out_proj", 7488, 768, 768), ("fc2", 7488, 768, 3072), ("cross_q", 7208, 768, 768)]:
    a, b, bias = rn(M, K).bfloat16(), (rn(N, K) * 0.05).bfloat16(), rn(N)
    if tag == "fc2":
        x, u = rn(M, N), rn(N)
        stats = torch.rand(M, K // 64, 2, device="cuda") + 1
        cases.append((tag, lambda a=a, b=b, bias=bias, x=x, u=u, stats=stats, K=K: ops.gemm(a, b, x, bias=bias, residual=x, rownorm=(stats, u, K))))
    elif tag == "out_proj":
        out = torch.empty(M, N, device="cuda")
        cases.append((tag, lambda a=a, b=b, bias=bias, out=out: ops.gemm(a, b, out, bias=bias)))
    else:
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        cases.append((tag, lambda a=a, b=b, bias=bias, out=out: ops.gemm(a, b, out, bias=bias, alpha=0.088, alpha_cols=768)))
for (tag, M, N, K, res) in [("l3.conv1", 7200, 256, 1024, False), ("l3.conv3", 7200, 1024, 256, True), ("l2.conv3", 28800, 512, 128, True),
                            ("l1.conv3", 115200, 256, 64, True), ("l1.conv1", 115200, 64, 256, False)]:
    a, b, sc, bi = rn(M, K).bfloat16(), (rn(N, K) * 0.05).bfloat16(), rn(N), rn(N)
    idn = rn(M, N).bfloat16() if res else None
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    cases.append((tag, lambda a=a, b=b, sc=sc, bi=bi, idn=idn, out=out: ops.gemm(a, b, out, scale=sc, bias=bi, act=ops.ACT_RELU, residual=idn)))
for (tag, n, h, c) in [("l3.conv2", 8, 30, 256), ("l2.conv2", 8, 60, 128)]:
    x, w, sc, bi = rn(n, h, h, c).bfloat16(), (rn(c, 9 * c) * 0.05).bfloat16(), rn(c), rn(c)
    cases.append((tag, lambda x=x, w=w, sc=sc, bi=bi: ops.conv3x3_s1(x, w, sc, bi)))
torch.cuda.synchronize()
for rep in range(2):  # second repetition = warm L2
    for tag, fn in cases:
        for fam in fams:
            os.environ["SGF_GEMM_FAMILY"] = fam
            torch.cuda.nvtx.range_push(f"{tag}:{fam}:{rep}")
            fn()
            torch.cuda.nvtx.range_pop()
torch.cuda.synchronize()
print("order:", [(t, f) for r in range(2) for t, _ in cases for f in fams])
