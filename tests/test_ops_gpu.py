"""-m gpu: every C-ABI kernel against a plain torch fp32 reference of the same op."""
import math

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


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 768, 768), (7488, 2304, 768), (901, 3072, 768),
                                   (1000, 64, 576), (257, 256, 2304), (77, 15, 768), (130, 40, 128)])
def test_gemm_plain_and_bias(ops, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ b.float().t() + bias
    out = ops.gemm(a, b, bias=bias, out_dtype=torch.float32)
    assert _rel(out, ref) < 2e-5, _rel(out, ref)
    out16 = ops.gemm(a, b, bias=bias)
    assert out16.dtype == torch.bfloat16
    assert _rel(out16, ref) < 4e-3


def test_gemm_epilogues(ops):
    g = torch.Generator(device="cuda").manual_seed(7)
    M, N, K = 515, 768, 1024
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    scale = torch.rand(N, device="cuda", generator=g) + 0.5
    res16 = torch.randn(M, N, device="cuda", generator=g).bfloat16()
    res32 = torch.randn(M, N, device="cuda", generator=g)
    acc = a.float() @ b.float().t()
    # scale+bias+relu after residual (bottleneck conv3)
    out = ops.gemm(a, b, scale=scale, bias=bias, residual=res16, act=ops.ACT_RELU, out_dtype=torch.float32)
    assert _rel(out, F.relu(acc * scale + bias + res16.float())) < 2e-5
    # gelu (fc1)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU, out_dtype=torch.float32)
    assert _rel(out, F.gelu(acc + bias)) < 2e-5
    # fp32 residual stream (fc2)
    out = ops.gemm(a, b, bias=bias, residual=res32, out_dtype=torch.float32)
    assert _rel(out, acc + bias + res32) < 2e-5
    # q scaling on the first 256 columns
    out = ops.gemm(a, b, bias=bias, alpha=0.125, alpha_cols=256, out_dtype=torch.float32)
    ref = acc + bias
    ref[:, :256] *= 0.125
    assert _rel(out, ref) < 2e-5


def test_gemm_pair_kernel_gelu(ops):
    """bias + GELU with a bf16 output on a shape the persistent CTA-pair kernel takes (M >= 2048, N >= 2048)."""
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 2400, 2048, 256
    a = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    b = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device="cuda", generator=g)
    out = ops.gemm(a, b, bias=bias, act=ops.ACT_GELU)
    assert out.dtype == torch.bfloat16
    assert _rel(out, F.gelu(a.float() @ b.float().t() + bias)) < 4e-3


def test_gemm_batched_heads(ops):
    """abs-pos bias: per-head [T,64] x [T,64]^T out of [T, H*64] buffers, padded fp32 output rows."""
    g = torch.Generator(device="cuda").manual_seed(3)
    T, H, dh = 333, 12, 64
    pq = torch.randn(T, H * dh, device="cuda", generator=g).bfloat16()
    pk = torch.randn(T, H * dh, device="cuda", generator=g).bfloat16()
    Tp = (T + 63) // 64 * 64
    out = torch.zeros(H, T, Tp, device="cuda")
    ops.gemm(pq, pk, out, M=T, N=T, K=dh, batch=H, lda=H * dh, ldb=H * dh, a_batch_stride=dh, b_batch_stride=dh,
             ldc=Tp, c_batch_stride=T * Tp)
    ref = torch.einsum("ihd,jhd->hij", pq.float().view(T, H, dh), pk.float().view(T, H, dh))
    assert _rel(out[:, :, :T], ref) < 2e-5
    assert out[:, :, T:].abs().max().item() == 0.0


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 30, 30, 256, 256), (1, 120, 120, 64, 64), (2, 60, 60, 128, 128),
                                            (1, 8, 8, 64, 64), (1, 33, 17, 64, 128)])
def test_conv3x3(ops, n, h, w, cin, cout):
    g = torch.Generator(device="cuda").manual_seed(n * h + cin)
    x = torch.randn(n, h, w, cin, device="cuda", generator=g).bfloat16()
    wt = (torch.randn(cout, cin, 3, 3, device="cuda", generator=g) / math.sqrt(9 * cin)).bfloat16()
    scale = torch.rand(cout, device="cuda", generator=g) + 0.5
    bias = torch.randn(cout, device="cuda", generator=g)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float(), padding=1)
    ref = F.relu(ref * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)).permute(0, 2, 3, 1)
    out = ops.conv3x3_s1(x, wt.permute(0, 2, 3, 1).contiguous(), scale, bias, act=ops.ACT_RELU)
    assert _rel(out, ref) < 4e-3, _rel(out, ref)


def test_stem_helpers(ops):
    g = torch.Generator(device="cuda").manual_seed(11)
    x = torch.randn(2, 3, 64, 48, device="cuda", generator=g)
    y = ops.nchw_to_nhwc_bf16(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).bfloat16())
    # (the strided convolutions and conv1 are im2col-free: tests/test_gemm_ts_gpu.py::test_conv2d_strided_im2col_free,
    #  ::test_conv1_7x7_stride2_window_mode)
    x2 = torch.randn(2, 20, 20, 64, device="cuda", generator=g).bfloat16()
    # maxpool
    mp = ops.maxpool3x3s2(x2)
    ref = F.max_pool2d(x2.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1).bfloat16()
    assert torch.equal(mp, ref)


@pytest.mark.parametrize("D", [768, 3072, 1024, 256, 4096, 1280, 512, 5120, 1536])
def test_row_layernorm(ops, D):
    g = torch.Generator(device="cuda").manual_seed(D)
    rows = 301
    x = (torch.randn(rows, D, device="cuda", generator=g) * 3 + 1).bfloat16()
    g1, b1 = torch.rand(D, device="cuda", generator=g) + 0.5, torch.randn(D, device="cuda", generator=g)
    g2, b2 = torch.rand(D, device="cuda", generator=g) + 0.5, torch.randn(D, device="cuda", generator=g)
    out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    ops.row_layernorm(x, ln2=(g1, b1), out2=out)
    ref = F.layer_norm(x.float(), (D,), g1, b1, 1e-5)
    assert _rel(out, ref) < 4e-3
    assert (out.float() - ref).abs().max() <= 0.02 * ref.abs().max()
    # res + LN1(x) -> out1 (fp32 stream), LN2(out1) -> out2
    res = torch.randn(rows, D, device="cuda", generator=g)
    o1 = torch.empty(rows, D, device="cuda")
    o2 = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    ops.row_layernorm(x, ln1=(g1, b1), residual=res, out1=o1, ln2=(g2, b2), out2=o2)
    r1 = res + F.layer_norm(x.float(), (D,), g1, b1, 1e-5)
    assert _rel(o1, r1) < 1e-5
    assert _rel(o2, F.layer_norm(r1, (D,), g2, b2, 1e-5)) < 4e-3
    # a large common offset must not cost precision (centred variance)
    xo = (torch.randn(rows, D, device="cuda", generator=g) * 0.5 + 30).bfloat16()
    ops.row_layernorm(xo, ln2=(g1, b1), out2=out)
    assert _rel(out, F.layer_norm(xo.float(), (D,), g1, b1, 1e-5)) < 4e-3


def test_row_layernorm_gather_segments(ops):
    g = torch.Generator(device="cuda").manual_seed(5)
    D, V, B, Tt, P = 768, 1000, 3, 36, 64
    table = torch.randn(V, D, device="cuda", generator=g)
    tok = torch.randint(0, V, (B * Tt,), device="cuda", generator=g)
    type_emb = torch.randn(D, device="cuda", generator=g)
    gg, bb = torch.rand(D, device="cuda", generator=g) + 0.5, torch.randn(D, device="cuda", generator=g)
    zero = torch.zeros(B * Tt, dtype=torch.uint8, device="cuda")
    zero[5] = 1
    xbuf = torch.full((B, P + Tt, D), 7.0, device="cuda")
    ops.row_layernorm(table, rows=B * Tt, gather_idx=tok, pre_add=type_emb, ln1=(gg, bb), out1=xbuf.view(-1, D),
                      zero_row=zero, seg=(Tt, P + Tt, P))
    ref = F.layer_norm(table[tok] + type_emb, (D,), gg, bb, 1e-5).view(B, Tt, D)
    ref.view(-1, D)[5] = 0
    assert _rel(xbuf[:, P:], ref) < 1e-5
    assert (xbuf[:, :P] == 7.0).all()


def _attn_ref(q, k, v, bias, causal, kpm, head_scale):
    s = torch.einsum("bihd,bjhd->bhij", q.float(), k.float())
    if bias is not None:
        s = s + bias[None, :, : s.shape[2], : s.shape[3]]
    if causal:
        s = s + torch.full(s.shape[-2:], float("-inf"), device=s.device).triu(1)
    if kpm is not None:
        s = s.masked_fill(kpm[:, None, None, :].bool(), float("-inf"))
    p = s.softmax(-1)
    o = torch.einsum("bhij,bjhd->bihd", p, v.float())
    if head_scale is not None:
        o = o * head_scale.view(1, 1, -1, 1)
    return o


@pytest.mark.parametrize("B,H,Tq,Tk,causal,use_bias,use_kpm", [
    (2, 3, 200, 200, False, True, False),
    (1, 12, 936, 936, False, True, False),
    (2, 4, 133, 133, True, True, False),
    (1, 12, 901, 901, True, True, False),
    (2, 2, 130, 200, False, True, True),
    (1, 2, 64, 64, False, False, False),
    (1, 1, 65, 100, False, False, False),
])
def test_attention(ops, B, H, Tq, Tk, causal, use_bias, use_kpm):
    g = torch.Generator(device="cuda").manual_seed(B * 1000 + Tq + Tk)
    dh = 64
    D = H * dh
    # self-attention layout: one [B, T, 3D] buffer when Tq == Tk, else separate q and kv buffers
    q = torch.randn(B, Tq, H, dh, device="cuda", generator=g).bfloat16() * 0.35
    k = torch.randn(B, Tk, H, dh, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Tk, H, dh, device="cuda", generator=g).bfloat16()
    Tkp = (Tk + 63) // 64 * 64
    bias = None
    if use_bias:
        bias = torch.zeros(H, Tq, Tkp, device="cuda")
        bias[:, :, :Tk] = torch.randn(H, Tq, Tk, device="cuda", generator=g)
        bias = bias.half()  # the kernel streams the bias as fp16
    kpm = None
    if use_kpm:
        kpm = torch.zeros(B, Tk, dtype=torch.uint8, device="cuda")
        kpm[0, Tk - 37:] = 1
        kpm[1, 5] = 1
    hs = torch.rand(H, device="cuda", generator=g) + 0.5
    out = torch.empty(B, Tq, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(q, k, v, out, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D),
                  v_strides=(D, Tk * D), o_strides=(D, Tq * D), bias=bias, head_scale=hs, key_padding_mask=kpm,
                  causal=causal)
    ref = _attn_ref(q, k, v, bias.float() if bias is not None else None, causal, kpm, hs).reshape(B, Tq, D)
    err = _rel(out, ref)
    assert err < 8e-3, err


@pytest.mark.parametrize("Tq,Tk,causal,use_bias", [(200, 333, False, True), (260, 260, True, False), (130, 700, False, False)])
def test_attention_lazy_rescale_paths(ops, Tq, Tk, causal, use_bias):
    """Scores that keep GROWING along the keys (by far more than the lazy-rescaling threshold 2^8 per 32-key sub-tile, at a
    row-dependent rate, so that within a warp some rows move their reference maximum and others do not): exercises the rare
    paths of the online softmax -- the output accumulator rescaled in tensor memory, the first half of P(j) rescaled after
    the second half raised the reference, the running sum -- and the log-sum-exp side output, against fp32 softmax."""
    g = torch.Generator(device="cuda").manual_seed(Tq + Tk)
    B, H, dh = 2, 2, 64
    D = H * dh
    rate = torch.rand(B, Tq, H, 1, device="cuda", generator=g) * 1.5 + 0.05  # per-row growth rate
    q = (torch.randn(B, Tq, H, dh, device="cuda", generator=g) * 0.05 + rate / 8).bfloat16()
    ramp = torch.linspace(-4.0, 4.0, Tk, device="cuda").view(1, Tk, 1, 1)    # key scale: scores ~ 8 * rate * ramp * 8
    k = (torch.randn(B, Tk, H, dh, device="cuda", generator=g) * 0.05 + ramp).bfloat16()
    v = torch.randn(B, Tk, H, dh, device="cuda", generator=g).bfloat16()
    Tkp = (Tk + 63) // 64 * 64
    bias = None
    if use_bias:
        bias = torch.zeros(H, Tq, Tkp, device="cuda")
        bias[:, :, :Tk] = torch.randn(H, Tq, Tk, device="cuda", generator=g) * 3
        bias = bias.half()
    s = torch.einsum("bihd,bjhd->bhij", q.float(), k.float())
    assert (s.max(-1).values - s[..., :32].max(-1).values).max() > 40  # the reference maximum really has to move, many times
    out = torch.empty(B, Tq, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, Tq, device="cuda")
    ops.attention(q, k, v, out, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D),
                  v_strides=(D, Tk * D), o_strides=(D, Tq * D), bias=bias, causal=causal, lse=lse)
    ref = _attn_ref(q, k, v, bias.float() if bias is not None else None, causal, None, None).reshape(B, Tq, D)
    assert torch.isfinite(out.float()).all()
    assert _rel(out, ref) < 8e-3, _rel(out, ref)
    if bias is not None:
        s = s + bias.float()[None, :, :Tq, :Tk]
    if causal:
        s = s + torch.full(s.shape[-2:], float("-inf"), device=s.device).triu(1)
    ref_lse = torch.logsumexp(s, dim=-1) * 1.4426950408889634  # the kernel reports log2-domain log-sum-exp
    assert torch.allclose(lse, ref_lse, rtol=1e-4, atol=2e-3), (lse - ref_lse).abs().max()


def test_attention_fused_qkv_layout(ops):
    g = torch.Generator(device="cuda").manual_seed(99)
    B, T, H, dh = 2, 150, 12, 64
    D = H * dh
    qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
    out = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
    ops.attention(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, q_strides=(3 * D, T * 3 * D),
                  k_strides=(3 * D, T * 3 * D), v_strides=(3 * D, T * 3 * D), o_strides=(D, T * D))
    q, k, v = (t.reshape(B, T, H, dh) for t in qkv.split(D, dim=-1))
    ref = _attn_ref(q, k, v, None, False, None, None).reshape(B, T, D)
    assert _rel(out, ref) < 8e-3


@pytest.mark.parametrize("B,C,hp,wp,h,w", [(2, 15, 30, 30, 480, 480), (1, 150, 8, 8, 128, 128), (1, 171, 32, 32, 500, 375),
                                           (2, 15, 30, 40, 480, 640), (2, 150, 30, 30, 480, 480)])
def test_upsample_argmax_bit_exact(ops, B, C, hp, wp, h, w):
    """The fused upsample+argmax against the reference's own op on the same box: F.interpolate(bilinear,
    align_corners=False) on CUDA tensors (what mmseg.ops.resize executes, seg_criterion.py:240) followed by argmax.
    Half of the batch is ADVERSARIAL: class 1 is class 0 plus a perturbation of a few fp32 ulps, so the argmax depends on
    the last bit of the interpolation -- only the exact rounding sequence of ATen's kernel gives zero mismatches there.
    ATen's CPU kernels round differently from its CUDA kernel (and from each other: the channels-last kernel pre-multiplies
    the four weights), so against the CPU oracle near-ties within a few ulps may differ; that count is reported."""
    g = torch.Generator(device="cuda").manual_seed(C + h)
    logits = torch.randn(B, hp * wp + 1, C, device="cuda", generator=g)
    logits[0, :, 1] = logits[0, :, 0] * (1 + 2e-7 * torch.randn(hp * wp + 1, device="cuda", generator=g))
    x = logits[:, :-1].reshape(B, hp, wp, C).permute(0, 3, 1, 2)
    ref_gpu = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).argmax(-1)
    ref_gpu_contig = F.interpolate(x.contiguous(), size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).argmax(-1)
    up_cpu = F.interpolate(x.cpu(), size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    ref_cpu = up_cpu.argmax(-1)
    mism = {}
    for name, mode in [("plain", ops.LERP_PLAIN)] + [(f"fma{v}", 8 + v) for v in range(8)]:
        m = ops.upsample_argmax(logits, hp, wp, h, w, arith=mode)
        mism[name] = dict(vs_aten_cuda=(m != ref_gpu).sum().item(), vs_aten_cuda_contiguous=(m != ref_gpu_contig).sum().item(),
                          vs_aten_cpu=(m.cpu() != ref_cpu).sum().item())
    print(f"upsample+argmax mismatches of {ref_gpu.numel()} pixels ({h * w} adversarial): {mism}")
    import json, os
    with open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "upsample_arith.jsonl"), "a") as f:
        f.write(json.dumps(dict(shape=[B, C, hp, wp, h, w], mismatches=mism)) + "\n")
    mask = ops.upsample_argmax(logits, hp, wp, h, w)  # default = ATen CUDA arithmetic for this channel count / layout
    assert mism["fma0"]["vs_aten_cuda_contiguous"] == 0, mism   # SGF_LERP_ATEN_CUDA == ATen's NCHW kernel, near-ties included
    assert mism["fma1" if C >= 16 else "fma0"]["vs_aten_cuda"] == 0, mism  # ... and its NHWC kernel from 16 channels on
    assert torch.equal(mask, ref_gpu), mism                     # bit-exact against the reference's op on the reference's layout
    assert torch.equal(mask[1:], ref_gpu[1:]) and (mask[1:].cpu() != ref_cpu[1:]).sum().item() == 0  # separated logits: every path agrees
    # against the CPU oracle a pixel may differ only where its top-2 margin is within a few fp32 ulps
    top2 = up_cpu.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]) / top2[..., 0].abs().clamp_min(1e-6)
    assert not ((mask.cpu() != ref_cpu) & (margin > 1e-6)).any()
    # histograms
    target = torch.randint(-1, C + 1, (B, h, w), device="cuda", generator=g)
    mask2, areas = ops.upsample_argmax(logits, hp, wp, h, w, target=target)
    assert torch.equal(mask2, mask)
    valid = (target >= 0) & (target < C)
    pred, tgt = mask[valid], target[valid]
    inter = pred[pred == tgt]
    ai = torch.histc(inter.float(), bins=C, min=0, max=C - 1)
    ap = torch.histc(pred.float(), bins=C, min=0, max=C - 1)
    al = torch.histc(tgt.float(), bins=C, min=0, max=C - 1)
    assert torch.equal(areas[0], ai) and torch.equal(areas[1], ap) and torch.equal(areas[2], al)


def test_build_attn_bias(ops):
    g = torch.Generator(device="cuda").manual_seed(21)
    H, T, Tp = 4, 100, 128
    absb = torch.randn(H, T, Tp, device="cuda", generator=g)
    bucket = torch.randint(0, 50, (90, 90), device="cuda", generator=g)
    ids_a = torch.randint(0, 90, (64,), device="cuda", generator=g)
    ids_b = torch.randint(0, 90, (26,), device="cuda", generator=g)
    tab_a = torch.randn(50, H, device="cuda", generator=g)
    tab_b = torch.randn(50, H, device="cuda", generator=g)
    dense = torch.randn(H, T, Tp, device="cuda", generator=g)
    out = ops.build_attn_bias(absb, T, [(bucket, ids_a, tab_a, 0, 64), (bucket, ids_b, tab_b, 74, 100)], dense_add=dense)
    ref = (absb + dense)[:, :, :T].clone()
    ref[:, 0:64, 0:64] += tab_a[bucket[ids_a][:, ids_a]].permute(2, 0, 1)
    ref[:, 74:100, 74:100] += tab_b[bucket[ids_b][:, ids_b]].permute(2, 0, 1)
    assert torch.allclose(out[:, :, :T], ref, atol=1e-6)
    b16, b32 = ops.build_attn_bias(absb, T, [(bucket, ids_a, tab_a, 0, 64), (bucket, ids_b, tab_b, 74, 100)],
                                   dense_add=dense, f16=True, keep_f32=True)
    assert b16.dtype == torch.float16 and torch.equal(b16[:, :, :T], ref.half()) and (b16[:, :, T:] == 0).all()
    assert torch.equal(b32[:, :, :T], b16[:, :, :T].float())  # the adjoint kernels see exactly the forward's values


def test_embedding_bag_mean(ops):
    g = torch.Generator(device="cuda").manual_seed(31)
    V, D, B, P = 500, 768, 3, 16
    table = torch.randn(V, D, device="cuda", generator=g)
    lens = torch.randint(1, 5, (B, P), device="cuda", generator=g)
    L = int(lens.sum(1).max())
    tokens = torch.full((B, L), 1, dtype=torch.long, device="cuda")
    for b in range(B):
        n = int(lens[b].sum())
        tokens[b, :n] = torch.randint(4, V, (n,), device="cuda", generator=g)
    ends = lens.cumsum(1).reshape(-1)
    out = ops.embedding_bag_mean(tokens, ends, table, P)
    # reference: the flatten/offset arithmetic of encoder_module.py:529-538 + nn.EmbeddingBag(mean)
    flat = tokens[tokens != 1]
    off = ends.view(B, P)
    off = torch.cat([off.new_zeros(B, 1), off], 1)
    base = torch.cat([off.new_zeros(1), off[:-1, -1]]).cumsum(0)
    starts = (off + base.unsqueeze(1))[:, :-1].flatten()
    ref = F.embedding_bag(flat, table, starts, mode="mean")
    assert torch.allclose(out, ref, atol=1e-5)
    out16 = ops.embedding_bag_mean(tokens, ends, table.bfloat16(), P)
    assert _rel(out16, ref) < 5e-3


@pytest.mark.parametrize("C,hp,h,eps", [(15, 8, 128, 0.0), (150, 4, 64, 0.1), (171, 30, 480, 0.0)])
def test_upsample_ce_loss(ops, C, hp, h, eps):
    g = torch.Generator(device="cuda").manual_seed(C)
    B = 2
    logits = torch.randn(B, hp * hp + 1, C, device="cuda", generator=g) * 2
    target = torch.randint(-1, C + 1, (B, h, h), device="cuda", generator=g)
    loss, cnt = ops.upsample_ce_loss(logits, target, hp, hp, eps)
    x = logits[:, :-1].reshape(B, hp, hp, C).permute(0, 3, 1, 2)
    up = F.interpolate(x, size=(h, h), mode="bilinear", align_corners=False).permute(0, 2, 3, 1).reshape(-1, C)
    t = target.reshape(-1)
    valid = (t >= 0) & (t < C)
    ref = F.cross_entropy(up[valid], t[valid], label_smoothing=eps)
    assert int(cnt.item()) == int(valid.sum().item())
    assert abs(loss.item() - ref.item()) < 2e-4 * max(1.0, abs(ref.item())), (loss.item(), ref.item())


def test_gemm_folded_layernorm(ops):
    """ffn_layernorm folded into the fc1/fc2 epilogues == explicit LayerNorm between the GEMMs."""
    g = torch.Generator(device="cuda").manual_seed(77)
    M, D, Fd = 515, 768, 3072
    a = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    w1 = (torch.randn(Fd, D, device="cuda", generator=g) * 0.03).bfloat16()
    b1 = torch.randn(Fd, device="cuda", generator=g) * 0.1
    w2 = torch.randn(D, Fd, device="cuda", generator=g) * 0.02
    b2 = torch.randn(D, device="cuda", generator=g) * 0.1
    gam = 1 + 0.1 * torch.randn(Fd, device="cuda", generator=g)
    bet = 0.05 * torch.randn(Fd, device="cuda", generator=g)
    res = torch.randn(M, D, device="cuda", generator=g)
    stats = torch.full((M, Fd // 64, 2), 123.0, device="cuda")
    f = ops.gemm(a, w1, bias=b1, act=ops.ACT_GELU, rowstats_out=stats)
    ff = f.float()
    assert torch.allclose(stats[:, :, 0].sum(1), ff.sum(1), rtol=1e-4, atol=1e-2)
    assert torch.allclose(stats[:, :, 1].sum(1), (ff * ff).sum(1), rtol=1e-4, atol=1e-2)
    stats2 = torch.zeros_like(stats)
    f2 = ops.gemm(a, w1, bias=b1, act=ops.ACT_GELU, rowstats_out=stats2)
    assert torch.equal(stats, stats2) and torch.equal(f, f2)  # deterministic
    w2f = (w2 * gam).bfloat16()
    out = ops.gemm(f, w2f, bias=b2 + w2 @ bet, residual=res, rownorm=(stats, w2f.float().sum(1).contiguous(), Fd),
                   out_dtype=torch.float32)
    ref = F.layer_norm(ff, (Fd,), gam, bet, 1e-5) @ w2.t() + b2 + res
    assert _rel(out, ref) < 3e-3, _rel(out, ref)


def test_attention_full_size_repeatable(ops):
    """BASELINE cfg-2 encoder shape (B=8, H=12, T=936, bias): many co-resident CTAs exercise the
    mbarrier hand-offs under load; results must be identical across launches and match fp32."""
    g = torch.Generator(device="cuda").manual_seed(2024)
    B, H, T, dh = 8, 12, 936, 64
    D = H * dh
    qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
    bias = torch.zeros(H, T, 960, device="cuda")
    bias[:, :, :T] = torch.randn(H, T, T, device="cuda", generator=g)
    bias = bias.half()
    hs = torch.rand(H, device="cuda", generator=g) + 0.5
    outs = []
    for _ in range(6):
        out = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
        ops.attention(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, q_strides=(3 * D, T * 3 * D),
                      k_strides=(3 * D, T * 3 * D), v_strides=(3 * D, T * 3 * D), o_strides=(D, T * D), bias=bias,
                      head_scale=hs)
        outs.append(out)
    torch.cuda.synchronize()
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    q, k, v = (t.reshape(B, T, H, dh) for t in qkv.split(D, dim=-1))
    ref = _attn_ref(q[:2], k[:2], v[:2], bias.float(), False, None, hs).reshape(2, T, D)
    assert _rel(outs[0][:2], ref) < 8e-3


@pytest.mark.parametrize("B,P,C,k,iters", [(2, 900, 150, 3, 25), (1, 64, 15, 3, 4), (3, 100, 171, 5, 1), (1, 16, 15, 1, 0)])
def test_label_propagation(ops, B, P, C, k, iters):
    """seg_criterion.py:197-213: normalise -> cosine similarity -> top-k -> iters x gather-mean."""
    g = torch.Generator(device="cuda").manual_seed(P + C)
    feats = torch.randn(B, P, 1024, device="cuda", generator=g).relu().bfloat16()
    logits = torch.randn(B, P + 1, C, device="cuda", generator=g) * 2
    prob, nbr = ops.label_propagation(feats, logits, topk=k, iters=iters, temperature=0.7)
    fn = F.normalize(feats.float(), dim=-1)
    sim = fn @ fn.transpose(-1, -2)
    # every chosen neighbour is within bf16 rounding of the true k-th best similarity; self is the best
    vals = sim.gather(-1, nbr.long())
    kth = sim.topk(k, dim=-1).values
    assert (vals >= kth[..., -1:] - 2e-2).all()
    assert (nbr[..., 0].long() == torch.arange(P, device="cuda")).float().mean() > 0.99
    assert (vals[..., :-1] >= vals[..., 1:] - 2e-2).all()  # largest first
    # the propagation itself, on OUR neighbours, against the reference's indexing loop
    ref = (logits[:, :P] / 0.7).softmax(-1)
    bi = torch.arange(B, device="cuda").view(B, 1, 1).expand(B, P, k)
    for _ in range(iters):
        ref = ref[bi, nbr.long()].mean(dim=-2)
    assert prob.shape == (B, P, C)
    assert torch.allclose(prob, ref, atol=2e-6, rtol=1e-4)


def test_row_topk_ties_and_order(ops):
    from ifseg_b200 import _lib
    import ctypes as C

    x = torch.tensor([[1.0, 5.0, 5.0, 3.0, 5.0, -1.0, 2.0], [0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 9.0]], device="cuda")
    out = torch.empty(2, 3, dtype=torch.int32, device="cuda")
    lib = _lib.load()
    _lib.check(lib.sgf_row_topk(C.c_void_p(x.data_ptr()), 7, 2, 7, 3, C.c_void_p(out.data_ptr()),
                                C.c_void_p(torch.cuda.current_stream().cuda_stream)), "sgf_row_topk")
    assert out.tolist() == [[1, 2, 4], [6, 0, 1]]
