"""-m gpu: the persistent CTA-pair GEMM with the TMA-store epilogue (csrc/gemm_ts.cu), every tile width and every
epilogue of the segofa path, against torch fp32 and bit for bit against the one-tile-per-CTA kernel."""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


@pytest.fixture(scope="module")
def ops(cuda_device):
    from ifseg_b200 import ops as o

    return o


class _family:
    def __init__(self, fam, bn=None):
        self.fam, self.bn = fam, bn

    def __enter__(self):
        os.environ["SGF_GEMM_FAMILY"] = self.fam
        if self.bn:
            os.environ["SGF_GEMM_TS_BN"] = str(self.bn)

    def __exit__(self, *exc):
        os.environ.pop("SGF_GEMM_FAMILY", None)
        os.environ.pop("SGF_GEMM_TS_BN", None)


def _both(ops, bn, fn):
    """fn() under the TMA-store kernel at tile width bn and under the tile kernel; returns both outputs."""
    lib_before = ops._lib.load().sgf_launch_count()
    with _family("ts", bn):
        a = fn()
    assert ops._lib.load().sgf_launch_count() > lib_before
    with _family("tile"):
        b = fn()
    return a, b


# M covers: odd number of 128-row tiles (phantom half of the last pair), a ragged last tile, many rounds per cluster
@pytest.mark.parametrize("bn", [64, 128, 192, 256, 384])
@pytest.mark.parametrize("M", [515, 128, 7208, 33000])
def test_ts_epilogues(ops, bn, M):
    g = torch.Generator(device="cuda").manual_seed(bn + M)
    N, K = 768, 320 if M > 20000 else 1024
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    res16 = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    res32 = torch.randn(M, N, device="cuda", generator=g)
    acc = a.float() @ b.float().t()

    # plain bf16 / fp32 stores (dX, d(encoder_out))
    o, t = _both(ops, bn, lambda: ops.gemm(a, b))
    assert _rel(o, acc) < 4e-3 and torch.equal(o, t)
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, out_dtype=torch.float32))
    assert _rel(o, acc) < 2e-5 and torch.equal(o, t)
    # bias (+ fp32 out): out_proj, image_proj, position projections
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, bias=bias, out_dtype=torch.float32))
    assert _rel(o, acc + bias) < 2e-5 and torch.equal(o, t)
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, bias=bias))
    assert torch.equal(o, t)
    # q scaling on the first 256 columns (fused QKV)
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, bias=bias, alpha=0.125, alpha_cols=256))
    ref = acc + bias
    ref[:, :256] *= 0.125
    assert _rel(o, ref) < 4e-3 and torch.equal(o, t)
    # GELU (fc1), with and without the row statistics of the stored bf16 output
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, bias=bias, act=ops.ACT_GELU))
    assert _rel(o, F.gelu(acc + bias)) < 4e-3 and torch.equal(o, t)
    st_a = torch.full((M, N // 64, 2), 7.0, device="cuda")
    st_b = torch.full((M, N // 64, 2), 9.0, device="cuda")
    with _family("ts", bn):
        o = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, rowstats_out=st_a)
    with _family("tile"):
        t = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, rowstats_out=st_b)
    assert torch.equal(o, t)
    of = o.float().view(M, N // 64, 64)
    assert torch.allclose(st_a[..., 0], of.sum(-1), rtol=1e-5, atol=1e-4)
    assert torch.allclose(st_a[..., 1], (of * of).sum(-1), rtol=1e-5, atol=1e-4)
    # stem: BN affine + ReLU, + bf16 residual before the ReLU, affine only
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, scale=scale, bias=bias, act=ops.ACT_RELU))
    assert _rel(o, F.relu(acc * scale + bias)) < 4e-3 and torch.equal(o, t)
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, scale=scale, bias=bias, residual=res16, act=ops.ACT_RELU))
    assert _rel(o, F.relu(acc * scale + bias + res16.float())) < 4e-3 and torch.equal(o, t)
    o, t = _both(ops, bn, lambda: ops.gemm(a, b, scale=scale, bias=bias))
    assert torch.equal(o, t)

    # fp32 residual stream updated in place (fc2): TMA reduce-add
    def inplace():
        x = res32.clone()
        ops.gemm(a, b, x, bias=bias, residual=x)
        return x

    o, t = _both(ops, bn, inplace)
    assert _rel(o, acc + bias + res32) < 2e-5 and torch.equal(o, t)


@pytest.mark.parametrize("bn", [64, 256, 384])
def test_ts_folded_layernorm(ops, bn):
    """fc1 -> ffn_layernorm -> fc2 with the LayerNorm folded into the two epilogues, x updated in place."""
    g = torch.Generator(device="cuda").manual_seed(77)
    M, D, Fd = 1301, 768, 3072
    a = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(Fd, D, device="cuda", generator=g) * 0.03).bfloat16()
    b1 = torch.randn(Fd, device="cuda", generator=g) * 0.1
    w2 = torch.randn(D, Fd, device="cuda", generator=g) * 0.02
    b2 = torch.randn(D, device="cuda", generator=g) * 0.1
    gam = 1 + 0.1 * torch.randn(Fd, device="cuda", generator=g)
    bet = 0.05 * torch.randn(Fd, device="cuda", generator=g)
    res = torch.randn(M, D, device="cuda", generator=g)
    w2f = (w2 * gam).bfloat16()
    u = w2f.float().sum(1).contiguous()

    def run():
        stats = torch.zeros((M, Fd // 64, 2), device="cuda")
        f = ops.gemm(a, w1, bias=b1, act=ops.ACT_GELU, rowstats_out=stats)
        x = res.clone()
        ops.gemm(f, w2f, x, bias=b2 + w2 @ bet, residual=x, rownorm=(stats, u, Fd))
        return f, x

    with _family("ts", bn):
        f, x = run()
    with _family("tile"):
        f_t, x_t = run()
    ref = F.layer_norm(f.float(), (Fd,), gam, bet, 1e-5) @ w2.t() + b2 + res
    assert torch.equal(f, f_t)
    assert _rel(x, ref) < 3e-3, _rel(x, ref)
    assert _rel(x, x_t) < 1e-6


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 30, 30, 256, 256), (1, 120, 120, 64, 64), (2, 60, 60, 128, 128),
                                            (1, 8, 8, 64, 64), (1, 33, 17, 64, 128), (3, 30, 30, 128, 192)])
def test_ts_conv3x3(ops, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(n * h + cin)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)).bfloat16()
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    wk = wt.permute(0, 2, 3, 1).contiguous()
    for bn in (64, 128, 192, 256):
        if cout % bn:
            continue
        with _family("ts", bn):
            out = ops.conv3x3_s1(x, wk, scale, bias, act=ops.ACT_RELU)
        with _family("tile"):
            out_t = ops.conv3x3_s1(x, wk, scale, bias, act=ops.ACT_RELU)
        assert _rel(out, ref) < 4e-3, (bn, _rel(out, ref))
        assert torch.equal(out, out_t), bn


@pytest.mark.parametrize("n,h,w,cin,cout,k,stride,pad", [
    (2, 60, 60, 128, 128, 3, 2, 1),    # layer2 block 0 conv2 (3x3/2)
    (2, 120, 120, 256, 512, 1, 2, 0),  # layer2 downsample (1x1/2)
    (1, 60, 60, 512, 1024, 1, 2, 0),   # layer3 downsample
    (2, 30, 30, 256, 256, 3, 1, 1),    # 3x3/1 through the general entry
    (1, 33, 17, 64, 128, 3, 2, 1),     # odd extents, ragged tiles
    (3, 31, 45, 64, 64, 1, 2, 0),
])
def test_conv2d_strided_im2col_free(ops, n, h, w, cin, cout, k, stride, pad):
    """Strided / padded convolutions read their input through TMA element strides: against F.conv2d."""
    g = torch.Generator(device="cuda").manual_seed(n * h + cin + k)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, k, k, device="cuda", generator=g) / math.sqrt(k * k * cin)).bfloat16()
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), stride=stride, padding=pad)
    ref = ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)
    wk = wt.permute(0, 2, 3, 1).reshape(cout, -1).contiguous()
    for act in (ops.ACT_RELU, ops.ACT_NONE):
        out = ops.conv2d(x, wk, scale, bias, kh=k, kw=k, stride=stride, pad=pad, act=act)
        exp = (F.relu(ref) if act == ops.ACT_RELU else ref).permute(0, 2, 3, 1)
        assert out.shape == exp.shape
        assert _rel(out, exp) < 4e-3, _rel(out, exp)
        # one storage rounding away from the fp32 result: elementwise within 1 bf16 ulp (+ accumulation noise)
        assert (out.float() - exp).abs().max() <= 2e-2 * exp.abs().max()


@pytest.mark.parametrize("n,H,W", [(2, 64, 48), (1, 480, 480), (2, 130, 70)])
def test_conv1_7x7_stride2_window_mode(ops, n, H, W):
    """conv1 of the stem: [N,3,H,W] fp32 -> zero-padded NHWC8 -> 7 filter-row taps of overlapping 128-byte windows."""
    g = torch.Generator(device="cuda").manual_seed(H + W)
    x = torch.randn(n, 3, H, W, device="cuda", generator=g)
    wt = (torch.randn(64, 3, 7, 7, device="cuda", generator=g) / 12).bfloat16()
    scale = torch.rand(64, device="cuda", generator=g) + 0.5
    bias = torch.randn(64, device="cuda", generator=g)
    pad = 3
    hp, wp = H + 2 * pad, (W + 2 * pad + 8 + 7) // 8 * 8
    xp = ops.nchw_to_nhwc8_padded(x, pad, hp, wp)
    assert xp.shape == (n, hp, wp, 8)
    assert torch.equal(xp[:, pad:pad + H, pad:pad + W, :3], x.permute(0, 2, 3, 1).bfloat16())
    assert xp[..., 3:].abs().max() == 0 and xp[:, :pad].abs().max() == 0 and xp[:, :, :pad].abs().max() == 0
    assert xp[:, pad + H:].abs().max() == 0 and xp[:, :, pad + W:].abs().max() == 0
    wm = torch.zeros(64, 7, 8, 8, device="cuda")
    wm[:, :, :7, :3] = wt.float().permute(0, 2, 3, 1)
    out = ops.conv2d(xp, wm.reshape(64, 448).bfloat16().contiguous(), scale, bias, kh=7, kw=1, stride=2, act=ops.ACT_RELU,
                     window=(H, W, pad, 3, 7))
    ref = F.conv2d(x.bfloat16().float(), wt.float(), stride=2, padding=3)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    assert out.shape == ref.shape
    assert _rel(out, ref) < 4e-3, _rel(out, ref)
