"""-m gpu: the adjoint (training-path) kernels against torch autograd of the same op in fp32."""
import math

import pytest
import torch
import torch.nn.functional as F

from test_ops_gpu import _attn_ref, _rel

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops(cuda_device):
    from ifseg_b200 import ops as o

    return o


def _gen(seed):
    return torch.Generator(device="cuda").manual_seed(seed)


# ----------------------------------------------------------------------------------------
# row kernel: forward x_act, and the adjoint in the three site shapes the engine uses
# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("D", [3072, 4096, 256, 1024, 2048, 5120, 1536])
def test_row_layernorm_gelu_forward(ops, D):
    g = _gen(D)
    rows = 77
    h = torch.randn(rows, D, device="cuda", generator=g).bfloat16()
    gam, bet = torch.rand(D, device="cuda", generator=g) + 0.5, torch.randn(D, device="cuda", generator=g)
    out = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    ops.row_layernorm(h, ln2=(gam, bet), out2=out, x_act=ops.ACT_GELU)
    ref = F.layer_norm(F.gelu(h.float()), (D,), gam, bet, 1e-5)
    assert _rel(out, ref) < 4e-3


@pytest.mark.parametrize("D,rows", [(768, 901), (1024, 130), (256, 9), (1280, 33), (512, 70)])
def test_row_layernorm_bwd_residual_site(ops, D, rows):
    """y -> LN1 -> + residual -> out1 ; out2 = LN2(out1)   (attn_ln / final_layer_norm site)."""
    g = _gen(D + rows)
    y = torch.randn(rows, D, device="cuda", generator=g, requires_grad=True)
    res = torch.randn(rows, D, device="cuda", generator=g, requires_grad=True)
    g1, b1, g2, b2 = [(torch.rand(D, device="cuda", generator=g) + 0.5).requires_grad_() for _ in range(4)]
    dy2 = (torch.randn(rows, D, device="cuda", generator=g) * 0.1).bfloat16()
    dv_in = torch.randn(rows, D, device="cuda", generator=g) * 0.1
    v = F.layer_norm(y, (D,), g1, b1, 1e-5) + res
    out2 = F.layer_norm(v, (D,), g2, b2, 1e-5)
    (out2 * dy2.float()).sum().backward(retain_graph=True)
    (v * dv_in).sum().backward()
    d_res = torch.empty(rows, D, device="cuda")
    dx = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    pg = [torch.zeros(D, device="cuda") for _ in range(4)]
    cs = torch.zeros(D, device="cuda")
    ops.row_layernorm_bwd(rows=rows, D=D, x=y.detach(), g1=g1.detach(), v=v.detach(), g2=g2.detach(), dy2=dy2,
                          dv_in=dv_in, d_res=d_res, dx=dx, dg1=pg[0], db1=pg[1], dg2=pg[2], db2=pg[3], dx_colsum=cs)
    assert _rel(cs, y.grad.sum(0)) < 1e-4
    assert _rel(d_res, res.grad) < 1e-5
    assert _rel(dx, y.grad) < 4e-3
    for got, ref in zip(pg, (g1.grad, b1.grad, g2.grad, b2.grad)):
        assert _rel(got, ref) < 1e-4, _rel(got, ref)


def test_row_layernorm_bwd_in_place_and_plain_residual(ops):
    """x2 = x1 + y (no LN1), a = LN2(x2): d_res aliases dv_in, dx (= dy, bf16) is the same gradient."""
    g = _gen(5)
    rows, D = 333, 768
    v = torch.randn(rows, D, device="cuda", generator=g, requires_grad=True)
    g2, b2 = [(torch.rand(D, device="cuda", generator=g) + 0.5).requires_grad_() for _ in range(2)]
    dy2 = (torch.randn(rows, D, device="cuda", generator=g) * 0.1).bfloat16()
    dv = torch.randn(rows, D, device="cuda", generator=g) * 0.1
    (F.layer_norm(v, (D,), g2, b2, 1e-5) * dy2.float()).sum().backward(retain_graph=True)
    (v * dv).sum().backward()
    stream = dv.clone()
    dx = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    dg2, db2 = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ops.row_layernorm_bwd(rows=rows, D=D, v=v.detach(), g2=g2.detach(), dy2=dy2, dv_in=stream, d_res=stream, dx=dx,
                          dg2=dg2, db2=db2)
    assert _rel(stream, v.grad) < 1e-5
    assert _rel(dx, v.grad) < 4e-3
    assert _rel(dg2, g2.grad) < 1e-4 and _rel(db2, b2.grad) < 1e-4


@pytest.mark.parametrize("F_", [3072, 4096, 1024, 2048, 512, 5120, 1536])
def test_row_layernorm_bwd_gelu_site(ops, F_):
    """z = LN(gelu(h)): dh from dz (FFN site, no saved v: recomputed from h)."""
    g = _gen(F_)
    rows = 150
    h16 = torch.randn(rows, F_, device="cuda", generator=g).bfloat16()
    h = h16.float().requires_grad_()
    g2, b2 = [(torch.rand(F_, device="cuda", generator=g) + 0.5).requires_grad_() for _ in range(2)]
    dz = (torch.randn(rows, F_, device="cuda", generator=g) * 0.1).bfloat16()
    (F.layer_norm(F.gelu(h), (F_,), g2, b2, 1e-5) * dz.float()).sum().backward()
    dh = torch.empty(rows, F_, device="cuda", dtype=torch.bfloat16)
    dg2, db2, cs = torch.zeros(F_, device="cuda"), torch.zeros(F_, device="cuda"), torch.zeros(F_, device="cuda")
    ops.row_layernorm_bwd(rows=rows, D=F_, x=h16, x_act=ops.ACT_GELU, g2=g2.detach(), dy2=dz, dx=dh, dg2=dg2, db2=db2,
                          dx_colsum=cs)
    assert _rel(dh, h.grad) < 5e-3, _rel(dh, h.grad)
    assert _rel(dg2, g2.grad) < 1e-4 and _rel(db2, b2.grad) < 1e-4
    assert _rel(cs, h.grad.sum(0)) < 2e-3, _rel(cs, h.grad.sum(0))


def test_row_layernorm_bwd_embedding_site_scatter(ops):
    """gathered rows + pre_add -> LN1 -> out1 (segment-mapped) ; LN2: d_pre_add, and dx scattered/accumulated
    back to the gathered source rows (decoder input = encoder_out rows)."""
    g = _gen(11)
    B, P, Te, D = 2, 9, 14, 256
    Td = P + 1
    src = torch.randn(B * Te, D, device="cuda", generator=g).bfloat16()
    srcf = src.float().requires_grad_()
    idx = (torch.arange(B).unsqueeze(1) * Te + torch.arange(P).unsqueeze(0)).reshape(-1).cuda()
    pre = torch.randn(D, device="cuda", generator=g).requires_grad_()
    g1, b1, g2, b2 = [(torch.rand(D, device="cuda", generator=g) + 0.5).requires_grad_() for _ in range(4)]
    dy2 = (torch.randn(B * Td, D, device="cuda", generator=g) * 0.1).bfloat16()
    dv_in = torch.randn(B * Td, D, device="cuda", generator=g) * 0.1
    t = srcf[idx] + pre
    v = F.layer_norm(t, (D,), g1, b1, 1e-5)  # rows r -> out rows (r // P) * Td + 1 + r % P
    rowmap = ((torch.arange(B * P) // P) * Td + 1 + torch.arange(B * P) % P).cuda()
    out2 = F.layer_norm(v, (D,), g2, b2, 1e-5)
    ((out2 * dy2.float()[rowmap]).sum() + (v * dv_in[rowmap]).sum()).backward()
    vbuf = torch.zeros(B * Td, D, device="cuda")
    vbuf[rowmap] = v.detach()
    base = torch.randn(B * Te, D, device="cuda", generator=g)
    dx = base.clone()
    pg = [torch.zeros(D, device="cuda") for _ in range(5)]
    ops.row_layernorm_bwd(rows=B * P, D=D, x=src, gather_idx=idx, pre_add=pre.detach(), g1=g1.detach(), v=vbuf,
                          g2=g2.detach(), dy2=dy2, dv_in=dv_in, dx=dx, dx_accumulate=True, dg1=pg[0], db1=pg[1],
                          dg2=pg[2], db2=pg[3], d_pre_add=pg[4], seg=(P, Td, 1))
    assert _rel(dx - base, srcf.grad) < 1e-4
    for got, ref in zip(pg, (g1.grad, b1.grad, g2.grad, b2.grad, pre.grad)):
        assert _rel(got, ref) < 1e-4, _rel(got, ref)


# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,dt", [(8920, 768, torch.bfloat16), (901, 3072, torch.bfloat16), (768, 2304, torch.float32),
                                    (150, 768, torch.float32), (77, 150, torch.bfloat16)])
def test_transpose_cast(ops, M, N, dt):
    g = _gen(M + N)
    ld = (N + 7) // 8 * 8
    buf = torch.randn(M, ld, device="cuda", generator=g).to(dt)
    x = buf[:, :N]
    colsum = torch.zeros(N, device="cuda")
    out_c = torch.empty(M, ld, device="cuda", dtype=torch.bfloat16)
    out_t = ops.transpose_cast(x, out_c=out_c, colsum=colsum)
    assert out_t.shape == (N, (M + 7) // 8 * 8)
    assert torch.equal(out_t[:, :M], x.bfloat16().t())
    assert (out_t[:, M:] == 0).all()
    assert torch.equal(out_c[:, :N], x.bfloat16())
    assert _rel(colsum, x.float().sum(0)) < 1e-5


def test_dense_adjoints_through_gemm(ops):
    """dX = dY W and dW = dY^T X with the forward GEMM on transposed operands."""
    g = _gen(3)
    M, K, N = 1115, 768, 3072
    x = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    w = (torch.randn(N, K, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    dy = (torch.randn(M, N, device="cuda", generator=g) * 0.1).bfloat16()
    wt = ops.transpose_cast(w)  # [K, N]
    dx = ops.gemm(dy, wt, M=M, N=K, K=N)
    assert _rel(dx, dy.float() @ w.float()) < 4e-3
    db = torch.zeros(N, device="cuda")
    dyt = ops.transpose_cast(dy, colsum=db)  # [N, pad8(M)]
    xt = ops.transpose_cast(x)  # [K, pad8(M)]
    dw = ops.gemm(dyt, xt, M=N, N=K, K=M, out_dtype=torch.float32)
    assert _rel(dw, dy.float().t() @ x.float()) < 2e-5
    assert _rel(db, dy.float().sum(0)) < 1e-5
    acc = torch.randn(N, K, device="cuda", generator=g)
    dw2 = ops.gemm(dyt, xt, acc.clone(), M=N, N=K, K=M, residual=acc)
    assert _rel(dw2, acc + dy.float().t() @ x.float()) < 2e-5


@pytest.mark.parametrize("T,N,K,split", [(8920, 768, 768, 8), (1115, 3072, 768, 2), (462, 2304, 768, 1), (901, 768, 3072, 3),
                                         (130, 64, 96, 1)])
def test_gemm_mixed_major_wgrad(ops, T, N, K, split):
    """dW[n,k] = sum_t dY[t,n] X[t,k] straight from the row-major dY / X (MN-major TMA operands), split-K atomics."""
    g = _gen(T + N + K)
    dy = (torch.randn(T, N, device="cuda", generator=g) * 0.1).bfloat16()
    x = torch.randn(T, K, device="cuda", generator=g).bfloat16()
    base = torch.randn(N, K, device="cuda", generator=g)
    out = base.clone()
    ops.gemm_ex(dy, x, out, M=N, N=K, K=T, a_mn=True, b_mn=True, split_k=split)
    if split == 1:
        ref = dy.float().t() @ x.float()
    else:
        ref = base + dy.float().t() @ x.float()
    assert _rel(out, ref) < 2e-5, _rel(out, ref)
    out16 = torch.empty(N, K, device="cuda", dtype=torch.bfloat16)
    ops.gemm_ex(dy, x, out16, M=N, N=K, K=T, a_mn=True, b_mn=True)
    assert _rel(out16, dy.float().t() @ x.float()) < 4e-3


@pytest.mark.parametrize("M,N,K", [(8920, 768, 3072), (462, 768, 2304), (901, 3072, 768)])
def test_gemm_mixed_major_dgrad(ops, M, N, K):
    """dX[m,n] = sum_k dY[m,k] W[k,n] with W as stored ([out=k, in=n], B MN-major)."""
    g = _gen(M + N + K)
    dy = (torch.randn(M, K, device="cuda", generator=g) * 0.1).bfloat16()
    w = (torch.randn(K, N, device="cuda", generator=g) / math.sqrt(K)).bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.gemm_ex(dy, w, out, M=M, N=N, K=K, a_mn=False, b_mn=True)
    assert _rel(out, dy.float() @ w.float()) < 4e-3


# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,hp,wp,h,w,eps", [(2, 15, 8, 8, 128, 128, 0.0), (1, 150, 4, 4, 64, 64, 0.1),
                                                (2, 171, 30, 30, 480, 480, 0.0), (1, 15, 6, 5, 100, 75, 0.0),
                                                (1, 15, 16, 12, 10, 9, 0.1), (1, 40, 5, 7, 83, 101, 0.0),
                                                (1, 8, 4, 4, 128, 128, 0.0)])
def test_upsample_ce_backward(ops, B, C, hp, wp, h, w, eps):
    g = _gen(C + hp)
    Td = hp * wp + 1
    logits = torch.randn(B, Td, C, device="cuda", generator=g) * 3
    target = torch.randint(-1, C, (B, h, w), device="cuda", generator=g)
    lg = logits.clone().requires_grad_()
    up = F.interpolate(lg[:, :-1].reshape(B, hp, wp, C).permute(0, 3, 1, 2), size=(h, w), mode="bilinear",
                       align_corners=False).permute(0, 2, 3, 1).reshape(-1, C)
    tg = target.reshape(-1)
    keep = tg >= 0
    loss_ref = F.cross_entropy(up[keep], tg[keep], label_smoothing=eps)
    loss_ref.backward()
    lse = torch.empty(B, h, w, device="cuda")
    acc = ops.upsample_ce_loss(logits, target, hp, wp, eps, lse_out=lse, raw=True)
    assert abs((acc[0] / acc[1]).item() - loss_ref.item()) < 1e-4 * max(1.0, abs(loss_ref.item()))
    ldc = (C + 7) // 8 * 8
    dl = torch.full((B, Td, ldc), 7.0, device="cuda", dtype=torch.bfloat16)
    ops.upsample_ce_loss_bwd(logits, target, lse, acc[1:], hp, wp, dl, eps, grad_scale=1.0)
    assert (dl[:, -1] == 0).all() and (dl[:, :, C:] == 0).all()
    assert _rel(dl[:, :, :C], lg.grad) < 5e-3, _rel(dl[:, :, :C], lg.grad)


# ----------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,H,Tq,Tk,causal,use_bias,use_kpm", [
    (1, 2, 64, 64, False, False, False),
    (2, 3, 200, 200, False, True, False),
    (2, 4, 133, 133, True, True, False),
    (1, 12, 901, 901, True, True, False),
    (1, 12, 901, 1115, False, True, False),
    (2, 2, 130, 200, False, True, True),
    (1, 1, 65, 100, False, False, False),
])
def test_attention_backward(ops, B, H, Tq, Tk, causal, use_bias, use_kpm):
    g = _gen(B * 1000 + Tq + Tk)
    dh = 64
    D = H * dh
    q = (torch.randn(B, Tq, H, dh, device="cuda", generator=g) * 0.35).bfloat16()
    k = torch.randn(B, Tk, H, dh, device="cuda", generator=g).bfloat16()
    v = torch.randn(B, Tk, H, dh, device="cuda", generator=g).bfloat16()
    Tkp = (Tk + 63) // 64 * 64
    bias = None
    if use_bias:
        bias = torch.zeros(H, Tq, Tkp, device="cuda")
        bias[:, :, :Tk] = torch.randn(H, Tq, Tk, device="cuda", generator=g)
        bias = bias.half().float()  # fp16-representable: the forward streams fp16, the adjoint kernels read fp32
    kpm = None
    if use_kpm:
        kpm = torch.zeros(B, Tk, dtype=torch.uint8, device="cuda")
        kpm[0, Tk - 37:] = 1
        kpm[1, 5] = 1
    hs = torch.rand(H, device="cuda", generator=g) + 0.5
    dout = (torch.randn(B, Tq, D, device="cuda", generator=g) * 0.2).bfloat16()

    qf, kf, vf, hsf = q.float().requires_grad_(), k.float().requires_grad_(), v.float().requires_grad_(), hs.clone().requires_grad_()
    ref = _attn_ref(qf, kf, vf, bias, causal, kpm, hsf).reshape(B, Tq, D)
    (ref * dout.float()).sum().backward()

    out = torch.empty(B, Tq, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, Tq, device="cuda")
    ops.attention(q, k, v, out, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D), k_strides=(D, Tk * D),
                  v_strides=(D, Tk * D), o_strides=(D, Tq * D), bias=bias.half() if bias is not None else None,
                  head_scale=hs, key_padding_mask=kpm, causal=causal, lse=lse)
    assert _rel(out, ref) < 8e-3
    dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
    delta = torch.empty(B, H, Tq, device="cuda")
    dhs = torch.zeros(H, device="cuda")
    ops.attention_bwd(q, k, v, out, dout, dq, dk, dv, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D),
                      k_strides=(D, Tk * D), v_strides=(D, Tk * D), o_strides=(D, Tq * D), do_strides=(D, Tq * D),
                      dq_strides=(D, Tq * D), dk_strides=(D, Tk * D), dv_strides=(D, Tk * D), lse=lse, delta=delta,
                      bias=bias.half() if bias is not None else None, head_scale=hs, d_head_scale=dhs,
                      key_padding_mask=kpm, causal=causal, dq_scale=1.0)
    if bias is not None:  # same adjoint through the key-major bias copy the dK/dV kernel prefers
        bt = ops.transpose_bias(bias.half(), Tk)
        assert torch.equal(bt[:, :Tk, :Tq], bias.half()[:, :, :Tk].transpose(1, 2)) and (bt[:, :, Tq:] == 0).all()
        dq2, dk2, dv2 = torch.empty_like(dq), torch.empty_like(dk), torch.empty_like(dv)
        ops.attention_bwd(q, k, v, out, dout, dq2, dk2, dv2, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=(D, Tq * D),
                          k_strides=(D, Tk * D), v_strides=(D, Tk * D), o_strides=(D, Tq * D), do_strides=(D, Tq * D),
                          dq_strides=(D, Tq * D), dk_strides=(D, Tk * D), dv_strides=(D, Tk * D), lse=lse,
                          delta=torch.empty_like(delta), bias=bias.half(), bias_t=bt, head_scale=hs,
                          key_padding_mask=kpm, causal=causal, dq_scale=1.0)
        assert torch.equal(dk2, dk) and torch.equal(dv2, dv) and torch.equal(dq2, dq)
    torch.cuda.synchronize()
    assert _rel(dv, vf.grad) < 1.2e-2, ("dv", _rel(dv, vf.grad))
    assert _rel(dq, qf.grad) < 1.5e-2, ("dq", _rel(dq, qf.grad))
    assert _rel(dk, kf.grad) < 1.5e-2, ("dk", _rel(dk, kf.grad))
    assert _rel(dhs, hsf.grad) < 4e-2, ("dhs", _rel(dhs, hsf.grad))


def test_attention_backward_fused_qkv_layout_and_scale(ops):
    g = _gen(99)
    B, T, H, dh = 2, 150, 12, 64
    D = H * dh
    qkv = (torch.randn(B, T, 3 * D, device="cuda", generator=g) * 0.5).bfloat16()
    dout = (torch.randn(B, T, D, device="cuda", generator=g) * 0.2).bfloat16()
    q, k, v = (t.reshape(B, T, H, dh).float().requires_grad_() for t in qkv.split(D, dim=-1))
    ref = _attn_ref(q, k, v, None, True, None, None).reshape(B, T, D)
    (ref * dout.float()).sum().backward()
    out = torch.empty(B, T, D, device="cuda", dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device="cuda")
    s3 = (3 * D, T * 3 * D)
    ops.attention(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, B=B, H=H, Tq=T, Tk=T, q_strides=s3, k_strides=s3,
                  v_strides=s3, o_strides=(D, T * D), causal=True, lse=lse)
    dqkv = torch.empty_like(qkv)
    delta = torch.empty(B, H, T, device="cuda")
    ops.attention_bwd(qkv, qkv[:, :, D:], qkv[:, :, 2 * D:], out, dout, dqkv, dqkv[:, :, D:], dqkv[:, :, 2 * D:], B=B,
                      H=H, Tq=T, Tk=T, q_strides=s3, k_strides=s3, v_strides=s3, o_strides=(D, T * D),
                      do_strides=(D, T * D), dq_strides=s3, dk_strides=s3, dv_strides=s3, lse=lse, delta=delta,
                      causal=True, dq_scale=0.5)
    gq, gk, gv = dqkv.split(D, dim=-1)
    assert _rel(gq, 0.5 * q.grad.reshape(B, T, D)) < 1.5e-2
    assert _rel(gk, k.grad.reshape(B, T, D)) < 1.5e-2
    assert _rel(gv, v.grad.reshape(B, T, D)) < 1.2e-2


# ----------------------------------------------------------------------------------------
def _fairseq_adam_reference(p, grad, m, v, step, lr, b1, b2, eps, wd):
    """Restatement of the reference optimizer, custom_fairseq/fairseq/optim/adam.py:214-235 (fp32 tensors, no amsgrad)."""
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    denom = v.sqrt().add_(eps)
    step_size = lr * (1 - b2 ** step) ** 0.5 / (1 - b1 ** step)
    if wd != 0:
        p.add_(p, alpha=-wd * lr)
    p.addcdiv_(m, denom, value=-step_size)


def test_adam_and_sumsq(ops):
    """sgf_adam_step against fairseq's Adam (denom = sqrt(v) + eps, NOT torch.optim.AdamW's sqrt(v)/sqrt(bc2) + eps:
    the two differ by the effective epsilon during the first ~3000 updates) -- small gradients make eps matter."""
    g = _gen(1)
    n = 100003
    n_pad = (n + 3) // 4 * 4
    p0 = torch.randn(n_pad, device="cuda", generator=g)
    grad = torch.randn(n_pad, device="cuda", generator=g) * 1e-7  # |g| ~ eps: the eps placement is visible
    ref = p0.double().clone()
    rm, rv = torch.zeros_like(ref), torch.zeros_like(ref)
    p, m, v = p0.clone(), torch.zeros_like(p0), torch.zeros_like(p0)
    scale = torch.tensor([0.5], device="cuda")
    for step in range(1, 4):
        _fairseq_adam_reference(ref, grad.double() * 0.5, rm, rv, step, 5e-3, 0.9, 0.999, 1e-8, 0.1)
        ops.adam_step(p, grad, m, v, lr=5e-3, weight_decay=0.1, step=step, grad_scale=scale)
    assert _rel(p - p0, (ref - p0.double()).float()) < 1e-4  # compare the UPDATE, not the (dominant) parameter value
    assert _rel(m, rm.float()) < 1e-6 and _rel(v, rv.float()) < 1e-6
    out = torch.zeros(1, device="cuda")
    ops.sumsq(grad[:n], out)
    assert abs(out.item() - grad[:n].double().pow(2).sum().item()) < 1e-3 * out.item()


def test_row_dropout_droppath_masks_forward_backward(ops):
    """Counter-based dropout / DropPath: mask values and rates, per-sample path masks, the adjoint regenerates the
    forward's mask exactly, and a different device-side step draws a different mask."""
    g = _gen(21)
    Bn, rps, D = 16, 64, 768
    rows = Bn * rps
    x = torch.ones(rows, D, device="cuda")
    res = torch.zeros(rows, D, device="cuda")
    one, zero = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    step = torch.tensor([3], dtype=torch.int32, device="cuda")
    drop = dict(p=0.25, path_p=0.5, seed=7, site=5, rows_per_sample=rps, step=step)
    out1 = torch.empty(rows, D, device="cuda")
    out2 = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
    ops.row_layernorm(x, residual=res, out1=out1, ln2=(one, zero), out2=out2, drop=drop)
    m = out1.clone()
    scale = 1.0 / (0.75 * 0.5)
    assert ((m == 0) | ((m - scale).abs() < 1e-3)).all()
    per_sample = m.view(Bn, rps * D)
    alive = per_sample.abs().sum(1) > 0
    assert 2 <= int(alive.sum()) <= 14                                    # DropPath: whole samples, p = 0.5
    assert (per_sample[~alive] == 0).all()
    keep = (per_sample[alive] > 0).float().mean().item()
    assert abs(keep - 0.75) < 0.01, keep                                  # element dropout inside surviving samples
    # the two 4-element halves of every 8-element chunk come from independent streams, whatever the key's parity
    for seed in (6, 7, 8, 9):
        ops.row_layernorm(x, residual=res, out1=out1, ln2=(one, zero), out2=out2,
                          drop=dict(p=0.25, path_p=0.0, seed=seed, site=5, rows_per_sample=rps, step=step))
        k8 = (out1 > 0).view(rows, D // 8, 8)
        same = (k8[..., :4] == k8[..., 4:]).float().mean().item()
        assert abs(same - (0.75 ** 2 + 0.25 ** 2)) < 0.01, (seed, same)  # iid: P(equal) = p^2 + (1-p)^2 = 0.625
    dy = torch.randn(rows, D, device="cuda", generator=g)
    dx = torch.empty(rows, D, device="cuda")
    ops.row_layernorm_bwd(rows=rows, D=D, dy2=dy, dx=dx, drop=drop)
    assert torch.equal(dx, dy * m)
    step += 1
    ops.row_layernorm(x, residual=res, out1=out1, ln2=(one, zero), out2=out2, drop=drop)
    assert not torch.equal(out1, m) and abs((out1 > 0).float().mean().item() - (m > 0).float().mean().item()) < 0.3
    # p = 0: bit-identical to the call without a drop description
    o_a, o_b = torch.empty_like(out1), torch.empty_like(out1)
    y = torch.randn(rows, D, device="cuda", generator=g)
    ops.row_layernorm(y, ln1=(one, zero), residual=res, out1=o_a, ln2=(one, zero), out2=out2)
    ops.row_layernorm(y, ln1=(one, zero), residual=res, out1=o_b, ln2=(one, zero), out2=out2,
                      drop=dict(p=0.0, path_p=0.0, seed=1, site=1, rows_per_sample=rps, step=step))
    assert torch.equal(o_a, o_b)


@pytest.mark.parametrize("B,C,hp,S", [(8, 150, 30, 480), (3, 15, 8, 128), (2, 171, 40, 640)])
def test_artificial_sample_generated_on_device(ops, B, C, hp, S):
    """segmentation_dataset.py:303-329 on the GPU: rebuild the expected bags / targets with torch from the label grids
    the kernel reports (nearest resize == F.interpolate(mode='nearest'), the torchvision Resize(NEAREST) the dataset uses)."""
    g = _gen(C + hp)
    Lmax = 4
    lens = torch.randint(1, Lmax + 1, (C,), device="cuda", generator=g).int()
    names = torch.randint(4, 50000, (C, Lmax), device="cuda", generator=g)
    step = torch.tensor([11], dtype=torch.int32, device="cuda")
    bag, ends, target, grid = ops.artificial_sample(names, lens, B, hp, S, seed=3, step=step, want_grid=True)
    P = hp * hp
    assert bag.shape == (B, P * Lmax) and ends.shape == (B * P,) and target.shape == (B, S * S + 1)
    shs = set()
    for b in range(B):
        sh, sw = int(grid[b, 0]), int(grid[b, 1])
        assert 1 <= sh <= 32 and 1 <= sw <= 32
        shs.add((sh, sw))
        lab = grid[b, 2: 2 + sh * sw].reshape(1, 1, sh, sw).float()
        assert lab.min() >= 0 and lab.max() < C
        low = F.interpolate(lab, size=(hp, hp), mode="nearest").long().reshape(-1)
        full = F.interpolate(lab, size=(S, S), mode="nearest").long().reshape(-1)
        assert torch.equal(target[b, :-1], full + 59457) and int(target[b, -1]) == 2
        l = lens[low].long()
        assert torch.equal(ends[b * P:(b + 1) * P], l.cumsum(0))
        exp = torch.cat([names[c, : int(n)] for c, n in zip(low.tolist(), l.tolist())])
        assert torch.equal(bag[b, : exp.numel()], exp) and (bag[b, exp.numel():] == 1).all()
    assert len(shs) > 1  # samples differ
    step += 1
    _, _, target2 = ops.artificial_sample(names, lens, B, hp, S, seed=3, step=step)
    assert not torch.equal(target, target2)  # a new step draws a new sample


def test_generated_sample_feeds_the_training_engine(cuda_device):
    """The generated tensors are exactly what aux_input / text2seg_target carry: one train step runs on them."""
    from ifseg_b200 import ops as o
    from ifseg_b200.seg_criterion import class_targets
    from ifseg_b200.segofa import SegOFAModel
    from ifseg_b200.synthetic import generate_state_dict, synthetic_inputs
    from ifseg_b200.train_engine import SegOFATrainEngine

    C, S, B = 15, 64, 2
    model = SegOFAModel.from_config("segofa_tiny", C, S)
    model.load_state_dict(generate_state_dict(model.cfg, 0), strict=True)
    model = model.cuda()
    eng = SegOFATrainEngine(model, stochastic=False)
    inp = synthetic_inputs(model.cfg, B, S, seed=1)
    g = _gen(5)
    lens = torch.randint(1, 4, (C,), device="cuda", generator=g).int()
    names = torch.randint(4, 50000, (C, 3), device="cuda", generator=g)
    bag, ends, t2s = o.artificial_sample(names, lens, B, S // 16, S, seed=9)
    aux = dict(src_tokens=inp["src_tokens"].cuda(), src_lengths=inp["src_lengths"].cuda(), patch_images=bag,
               patch_masks=ends, prev_output_tokens=inp["prev_output_tokens"].cuda())
    tgt = class_targets(t2s[:, :-1].reshape(B, S, S), 59457, C)
    loss, logits = eng.forward_backward(aux, tgt)
    assert torch.isfinite(loss) and logits.shape == (B, (S // 16) ** 2 + 1, C)
    assert torch.isfinite(eng.arena.grad32).all() and eng.arena.grad32.abs().sum() > 0
