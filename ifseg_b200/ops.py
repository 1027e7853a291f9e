"""Thin torch-tensor front-ends of the C ABI (one function per extern "C" entry point).

PyTorch is plumbing only here: it owns device memory (caching allocator) and the current
stream; every FLOP of the hot path runs in libsegofa_b200.so.  All wrappers launch on
torch.cuda.current_stream() so that DDP hooks, record_function ranges and CUDA-graph capture
see them.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import ACT_GELU, ACT_NONE, ACT_RELU, SGF_BF16, SGF_F32  # noqa: F401

_DT = {torch.bfloat16: SGF_BF16, torch.float32: SGF_F32}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _req(t, dtype=None, name="tensor"):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the segofa_b200 hot path has no CPU fallback")
    if dtype is not None and t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t


class KernelTimer:
    """Optional per-launch CUDA-event timing (used by bench.py to attribute the step to kernel
    families and to compute the roofline of the dominant one).  Off by default: zero overhead."""

    def __init__(self, fine=False):
        self.fine = fine  # split the GEMM family by call-site tag
        self.records = []  # (family, start_event, end_event, flops, bytes)

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for fam, s, e, fl, by in self.records:
            d = out.setdefault(fam, dict(launches=0, ms=0.0, flops=0.0, bytes=0.0))
            d["launches"] += 1
            d["ms"] += s.elapsed_time(e)
            d["flops"] += fl
            d["bytes"] += by
        return out


_TIMER = None


def set_timer(timer):
    global _TIMER
    _TIMER = timer


class _timed:
    def __init__(self, family, flops=0.0, nbytes=0.0):
        self.family, self.flops, self.nbytes = family, flops, nbytes

    def __enter__(self):
        if _TIMER is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *a):
        if _TIMER is not None:
            self.e.record()
            _TIMER.records.append((self.family, self.s, self.e, self.flops, self.nbytes))
        return False


# Optional launch tap (tests only): an object with before(kind, **inputs) -> token and after(token, out) called around every
# GEMM / convolution / attention / row-kernel launch, so a test can replay each launch of a real forward on the CPU from the
# launch's OWN inputs (tests/test_parity_gpu.py: single-storage-point parity).  None by default: zero overhead.
_TAP = None


def set_tap(tap):
    global _TAP
    _TAP = tap


def launch_count():
    return int(_lib.load().sgf_launch_count())


def reset_launch_count():
    _lib.load().sgf_reset_launch_count()


def gemm(a, b, out=None, *, bias=None, scale=None, residual=None, act=ACT_NONE, alpha=1.0, alpha_cols=0,
         out_dtype=torch.bfloat16, batch=1, M=None, N=None, K=None, lda=None, ldb=None, ldc=None, ldr=None,
         a_batch_stride=0, b_batch_stride=0, c_batch_stride=0, r_batch_stride=0, tag=None,
         rowstats_out=None, rownorm=None):
    """out[M,N] = epilogue(a[M,K] @ b[N,K]^T).  a, b bf16 with unit inner stride; explicit
    M/N/K/ld*/batch strides allow strided views (heads, concatenated buffers)."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    if M is None:
        M, K = a.shape[-2], a.shape[-1]
    if N is None:
        N = b.shape[-2]
    lda = a.stride(-2) if lda is None else lda
    ldb = b.stride(-2) if ldb is None else ldb
    if out is None:
        out = torch.empty((M, N) if batch == 1 else (batch, M, N), dtype=out_dtype, device=a.device)
        if batch > 1:
            c_batch_stride = M * N
    ldc = out.stride(-2) if ldc is None else ldc
    args = _lib.GemmArgs(
        _p(a), lda, a_batch_stride, _p(b), ldb, b_batch_stride, _p(out), ldc, c_batch_stride, _DT[out.dtype],
        M, N, K, batch, _p(scale), _p(bias), _p(residual),
        (residual.stride(-2) if ldr is None else ldr) if residual is not None else 0, r_batch_stride,
        _DT[residual.dtype] if residual is not None else SGF_BF16, act, float(alpha), int(alpha_cols),
        _p(rowstats_out), _p(rownorm[0]) if rownorm else None, _p(rownorm[1]) if rownorm else None,
        int(rownorm[2]) if rownorm else 0)
    tok = None
    if _TAP is not None:
        tok = _TAP.before("gemm", a=a, b=b, M=M, N=N, K=K, lda=lda, ldb=ldb, ldc=ldc, batch=batch, bias=bias, scale=scale,
                          residual=residual, ldr=(residual.stride(-2) if ldr is None else ldr) if residual is not None else 0,
                          act=act, alpha=alpha, alpha_cols=alpha_cols, a_batch_stride=a_batch_stride,
                          b_batch_stride=b_batch_stride, c_batch_stride=c_batch_stride, r_batch_stride=r_batch_stride,
                          rowstats_out=rowstats_out, rownorm=rownorm, tag=tag)
    with _timed("gemm_tcgen05" + (":" + tag if tag and _TIMER is not None and _TIMER.fine else ""), 2.0 * M * N * K * batch):
        _lib.check(lib.sgf_gemm_bf16(C.byref(args), _stream()), "sgf_gemm_bf16")
    if tok is not None:
        _TAP.after(tok, out)
    return out


def gemm_ex(a, b, out, *, M, N, K, a_mn, b_mn, lda=None, ldb=None, ldc=None, split_k=1, accumulate=False, tag=None):
    """out[M,N] (+)= A_eff @ B_eff^T with MN-major operands (see sgf_gemm_bf16_ex): a is [K,M] when a_mn else [M,K];
    b is [K,N] when b_mn else [N,K].  split_k > 1 accumulates into the fp32 `out`."""
    lib = _lib.load()
    _req(a, torch.bfloat16, "a")
    _req(b, torch.bfloat16, "b")
    lda = a.stride(-2) if lda is None else lda
    ldb = b.stride(-2) if ldb is None else ldb
    ldc = out.stride(-2) if ldc is None else ldc
    acc = accumulate or split_k > 1
    args = _lib.GemmArgs(_p(a), lda, 0, _p(b), ldb, 0, _p(out), ldc, 0, _DT[out.dtype], M, N, K, 1, None, None,
                         _p(out) if acc else None, ldc if acc else 0, 0, SGF_F32 if acc else SGF_BF16, ACT_NONE, 1.0, 0,
                         None, None, None, 0)
    with _timed("gemm_tcgen05" + (":" + tag if tag and _TIMER is not None and _TIMER.fine else ""), 2.0 * M * N * K):
        _lib.check(lib.sgf_gemm_bf16_ex(C.byref(args), 1 if a_mn else 0, 1 if b_mn else 0, int(split_k), _stream()),
                   "sgf_gemm_bf16_ex")
    return out


def conv3x3_s1(x, w, scale, bias, act=ACT_RELU, out=None, tag=None):
    """x [N,H,W,Cin] bf16 NHWC, w [Cout,3,3,Cin] bf16 -> [N,H,W,Cout] bf16."""
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    n, h, wd, cin = x.shape
    cout = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous()
    if out is None:
        out = torch.empty((n, h, wd, cout), dtype=torch.bfloat16, device=x.device)
    args = _lib.Conv3x3Args(_p(x), _p(w), _p(out), n, h, wd, cin, cout, _p(scale), _p(bias), act)
    tok = _TAP.before("conv3x3", x=x, w=w, scale=scale, bias=bias, act=act, tag=tag) if _TAP is not None else None
    with _timed("gemm_tcgen05" + (":" + tag if tag and _TIMER is not None and _TIMER.fine else ""), 2.0 * n * h * wd * cout * 9 * cin):
        _lib.check(lib.sgf_conv3x3_s1_nhwc(C.byref(args), _stream()), "sgf_conv3x3_s1_nhwc")
    if tok is not None:
        _TAP.after(tok, out)
    return out


def conv2d(x, w, scale, bias, *, kh, kw, stride=1, pad=0, act=ACT_RELU, out=None, tag=None, window=None):
    """im2col-free convolution + BN affine (+ ReLU) over bf16 NHWC (sgf_conv2d_nhwc).
    x [N,H,W,Cin] (Cin % 64 == 0), w [Cout, kh*kw*Cin] tap-major -> [N,Ho,Wo,Cout] bf16.
    window=(H, W, pad_phys, cin_real, kw_real): x is instead the zero-padded [N,Hp,Wp,8] image of the 7x7/2 stem convolution
    (ops.nchw_to_nhwc8_padded) read as overlapping 8-pixel x 8-channel windows (kh taps of 64 'channels', kw = 1)."""
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    _req(w, torch.bfloat16, "w")
    assert x.is_contiguous() and w.is_contiguous()
    cout = w.shape[0]
    if window is None:
        n, h, wd, cin = x.shape
        ho, wo = (h + 2 * pad - kh) // stride + 1, (wd + 2 * pad - kw) // stride + 1
        geo = (n, h, wd, cin, cin, cin * wd, cin * wd * h, ho, wo, kh, kw, stride, stride, pad, pad)
    else:
        H, W, pad_phys, _, kw_real = window
        n, hp, wp, c8 = x.shape
        assert c8 == 8 and kw == 1 and kw_real <= 8
        ho, wo = (H + 2 * pad_phys - kh) // stride + 1, (W + 2 * pad_phys - kw_real) // stride + 1
        assert (wo - 1) * stride + 8 <= wp and (ho - 1) * stride + kh <= hp
        # (c = 64-element window, w = window index with a pitch of `stride` pixels, h = padded row, n)
        geo = (n, hp, wo, 64, stride * 8, wp * 8, wp * 8 * hp, ho, wo, kh, 1, stride, 1, 0, 0)
    if out is None:
        out = torch.empty((n, geo[7], geo[8], cout), dtype=torch.bfloat16, device=x.device)
    args = _lib.Conv2dArgs(_p(x), geo[0], geo[1], geo[2], geo[3], geo[4], geo[5], geo[6], _p(w), _p(out), geo[7], geo[8], cout,
                           geo[9], geo[10], geo[11], geo[12], geo[13], geo[14], _p(scale), _p(bias), act)
    tok = _TAP.before("conv2d", x=x, w=w, scale=scale, bias=bias, act=act, kh=kh, kw=kw, stride=stride, pad=pad, window=window,
                      tag=tag) if _TAP is not None else None
    with _timed("gemm_tcgen05" + (":" + tag if tag and _TIMER is not None and _TIMER.fine else ""),
                2.0 * n * geo[7] * geo[8] * cout * (kh * kw * geo[3] if window is None else kh * window[4] * window[3])):
        _lib.check(lib.sgf_conv2d_nhwc(C.byref(args), _stream()), "sgf_conv2d_nhwc")
    if tok is not None:
        _TAP.after(tok, out)
    return out


def nchw_to_nhwc8_padded(x, pad, hp, wp):
    """[N,C<=8,H,W] fp32 -> zero-padded [N,hp,wp,8] bf16 with the image at (pad, pad)."""
    lib = _lib.load()
    _req(x, torch.float32, "x")
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty((n, hp, wp, 8), dtype=torch.bfloat16, device=x.device)
    with _timed("layout", nbytes=x.numel() * 4.0 + y.numel() * 2.0):
        _lib.check(lib.sgf_nchw_f32_to_nhwc8_padded(_p(x), _p(y), n, c, h, w, pad, hp, wp, _stream()),
                   "sgf_nchw_f32_to_nhwc8_padded")
    return y


def nchw_to_nhwc_bf16(x):
    lib = _lib.load()
    _req(x, torch.float32, "x")
    n, c, h, w = x.shape
    x = x.contiguous()
    y = torch.empty((n, h, w, c), dtype=torch.bfloat16, device=x.device)
    with _timed("layout", nbytes=x.numel() * 6.0):
        _lib.check(lib.sgf_nchw_f32_to_nhwc_bf16(_p(x), _p(y), n, c, h, w, _stream()), "sgf_nchw_f32_to_nhwc_bf16")
    return y


def maxpool3x3s2(x):
    lib = _lib.load()
    _req(x, torch.bfloat16, "x")
    n, h, w, c = x.shape
    ho, wo = (h + 2 - 3) // 2 + 1, (w + 2 - 3) // 2 + 1
    y = torch.empty((n, ho, wo, c), dtype=torch.bfloat16, device=x.device)
    with _timed("maxpool", nbytes=(x.numel() + y.numel()) * 2.0):
        _lib.check(lib.sgf_maxpool3x3s2_nhwc(_p(x), _p(y), n, h, w, c, ho, wo, _stream()), "sgf_maxpool3x3s2_nhwc")
    return y


def _drop_fields(drop):
    """drop = None | dict(p, path_p, seed, site, rows_per_sample, step (int32 device tensor or None))."""
    if not drop:
        return 0.0, 0.0, 0, 0, 0, None
    return (float(drop.get("p", 0.0)), float(drop.get("path_p", 0.0)), int(drop.get("seed", 0)) & 0xFFFFFFFF,
            int(drop.get("site", 0)), int(drop.get("rows_per_sample", 0)), _p(drop.get("step")))


def row_layernorm(x, *, rows=None, D=None, ldx=None, gather_idx=None, pre_add=None, ln1=None, residual=None,
                  ldr=None, out1=None, ld1=None, ln2=None, out2=None, ld2=None, zero_row=None, seg=None,
                  clear_rowstats=None, x_act=ACT_NONE, drop=None):
    """See sgf_row_layernorm in include/segofa_b200.h.  ln1/ln2 = (gamma, beta) fp32 tensors."""
    lib = _lib.load()
    _req(x, None, "x")
    rows = x.shape[0] if rows is None else rows
    D = x.shape[-1] if D is None else D
    seg_len, seg_stride, seg_off = seg if seg is not None else (0, 0, 0)
    args = _lib.RowLnArgs(
        _p(x), x.stride(-2) if ldx is None else ldx, _DT[x.dtype], _p(gather_idx), _p(pre_add),
        _p(ln1[0]) if ln1 else None, _p(ln1[1]) if ln1 else None,
        _p(residual), (residual.stride(-2) if ldr is None else ldr) if residual is not None else 0,
        _DT[residual.dtype] if residual is not None else SGF_BF16,
        _p(out1), (out1.stride(-2) if ld1 is None else ld1) if out1 is not None else 0,
        _DT[out1.dtype] if out1 is not None else SGF_BF16,
        _p(ln2[0]) if ln2 else None, _p(ln2[1]) if ln2 else None,
        _p(out2), (out2.stride(-2) if ld2 is None else ld2) if out2 is not None else 0,
        _p(zero_row), rows, D, seg_len, seg_stride, seg_off, _p(clear_rowstats), int(x_act), *_drop_fields(drop))
    nb = rows * D * (x.element_size() + (residual.element_size() if residual is not None else 0)
                     + (out1.element_size() if out1 is not None else 0) + (2 if out2 is not None else 0))
    tok = None
    if _TAP is not None:
        tok = _TAP.before("row_layernorm", x=x, rows=rows, D=D, ldx=x.stride(-2) if ldx is None else ldx, gather_idx=gather_idx,
                          pre_add=pre_add, ln1=ln1, residual=residual, ln2=ln2, zero_row=zero_row, seg=seg, x_act=x_act,
                          drop=drop)
    with _timed("row_layernorm" + (f":D{D}{'g1' if ln1 else ''}" if _TIMER is not None and _TIMER.fine else ""), nbytes=float(nb)):
        _lib.check(lib.sgf_row_layernorm(C.byref(args), _stream()), "sgf_row_layernorm")
    if tok is not None:
        _TAP.after(tok, (out1, out2))


def build_attn_bias(abs_bias, Tk, blocks=(), out=None, dense_add=None, f16=False, keep_f32=False):
    """out[h,i,j] = abs[h,i,j] (+ dense_add) + rel-pos lookups.  abs_bias fp32 [H,Tq,row_stride>=Tk];
    blocks: iterable of (bucket int64 2-D, ids int64 [hi-lo], table fp32 [num_rel,H], lo, hi).
    f16=True returns the fp16 tensor the attention kernel reads (zero padding columns); with keep_f32 also the fp32
    tensor holding exactly the same values (debugging / tests; the adjoint kernels read the fp16 tensor): (b16, b32)."""
    lib = _lib.load()
    _req(abs_bias, torch.float32, "abs_bias")
    out16 = torch.zeros(abs_bias.shape, dtype=torch.float16, device=abs_bias.device) if f16 else None
    if f16 and not keep_f32:
        out = None
    else:
        out = torch.empty_like(abs_bias) if out is None else out
        assert out.stride() == abs_bias.stride()
    assert abs_bias.stride(2) == 1
    H, Tq, _ = abs_bias.shape
    args = _lib.BiasArgs()
    args.out, args.abs = (out.data_ptr() if out is not None else None), abs_bias.data_ptr()
    args.out_f16 = out16.data_ptr() if out16 is not None else None
    args.head_stride, args.row_stride = abs_bias.stride(0), abs_bias.stride(1)
    args.dense_add = dense_add.data_ptr() if dense_add is not None else None
    args.H, args.Tq, args.Tk, args.num_blocks = H, Tq, Tk, len(blocks)
    for i, (bucket, ids, table, lo, hi) in enumerate(blocks):
        _req(bucket, torch.int64, "bucket")
        _req(ids, torch.int64, "ids")
        _req(table, torch.float32, "table")
        assert table.is_contiguous() and table.shape[1] == H and ids.numel() == hi - lo
        args.blocks[i] = _lib.RelBlock(bucket.data_ptr(), bucket.stride(0), ids.data_ptr(), table.data_ptr(), lo, hi)
    with _timed("attn_bias", nbytes=8.0 * H * Tq * Tk):
        _lib.check(lib.sgf_build_attn_bias(C.byref(args), _stream()), "sgf_build_attn_bias")
    if f16:
        return (out16, out) if keep_f32 else out16
    return out


def attention(q, k, v, out, *, B, H, Tq, Tk, q_strides, k_strides, v_strides, o_strides, bias=None,
              head_scale=None, key_padding_mask=None, causal=False, lse=None):
    """q/k/v/out: bf16 tensors used as base pointers; *_strides = (row_stride, batch_stride) in elements."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        _req(t, torch.bfloat16, n)
    if bias is not None:
        _req(bias, torch.float16, "bias")
    args = _lib.AttentionArgs(
        _p(q), q_strides[0], q_strides[1], _p(k), k_strides[0], k_strides[1], _p(v), v_strides[0], v_strides[1],
        _p(out), o_strides[0], o_strides[1],
        _p(bias), bias.stride(0) if bias is not None else 0, bias.stride(1) if bias is not None else 0,
        _p(head_scale), _p(key_padding_mask), B, H, Tq, Tk, 1 if causal else 0, _p(lse))
    pairs = Tq * Tk if not causal else Tq * (Tq + 1) / 2.0
    tok = None
    if _TAP is not None:
        tok = _TAP.before("attention", q=q, k=k, v=v, B=B, H=H, Tq=Tq, Tk=Tk, q_strides=q_strides, k_strides=k_strides,
                          v_strides=v_strides, o_strides=o_strides, bias=bias, head_scale=head_scale,
                          key_padding_mask=key_padding_mask, causal=causal)
    with _timed("attention_tcgen05", 4.0 * B * H * pairs * 64):
        _lib.check(lib.sgf_attention_bf16(C.byref(args), _stream()), "sgf_attention_bf16")
    if tok is not None:
        _TAP.after(tok, out)
    return out


LERP_DEFAULT, LERP_PLAIN, LERP_ATEN_CUDA, LERP_ATEN_CUDA_NHWC = -1, 0, 8, 9  # SGF_LERP_* (include/segofa_b200.h)


def upsample_argmax(logits, hp, wp, h, w, target=None, num_tokens=None, arith=LERP_DEFAULT):
    """logits fp32 [B, >=hp*wp, C] -> mask int64 [B,h,w] (+ optional (intersect, pred, label) areas).  `arith`: rounding
    sequence of the interpolation (default = ATen's CUDA kernel, the reference's execution path)."""
    lib = _lib.load()
    _req(logits, torch.float32, "logits")
    B, _, Cn = logits.shape
    assert logits.stride(2) == 1
    mask = torch.empty((B, h, w), dtype=torch.int64, device=logits.device)
    areas = None
    if target is not None:
        _req(target, torch.int64, "target")
        areas = torch.zeros((3, Cn), dtype=torch.float32, device=logits.device)
    args = _lib.SegmaskArgs(_p(logits), logits.stride(0), logits.stride(1), B, Cn, hp, wp, h, w, _p(mask),
                            _p(target), _p(areas[0]) if areas is not None else None,
                            _p(areas[1]) if areas is not None else None,
                            _p(areas[2]) if areas is not None else None, int(arith))
    with _timed("upsample_argmax", nbytes=float(mask.numel() * 8 + B * hp * wp * Cn * 4)):
        _lib.check(lib.sgf_upsample_argmax(C.byref(args), _stream()), "sgf_upsample_argmax")
    return (mask, areas) if target is not None else mask


def embedding_bag_mean(tokens, ends, table, P):
    """tokens int64 [B,L] (pads at the row tails), ends int64 [B*P] per-sample cumulative bag lengths,
    table [V,D] fp32/bf16 -> fp32 [B*P, D] bag means (nn.EmbeddingBag(mode='mean') of the image-free branch)."""
    lib = _lib.load()
    _req(tokens, torch.int64, "tokens")
    _req(ends, torch.int64, "ends")
    _req(table, None, "table")
    B = tokens.shape[0]
    D = table.shape[1]
    out = torch.empty((B * P, D), dtype=torch.float32, device=tokens.device)
    with _timed("embedding_bag", nbytes=float(out.numel() * 4)):
        _lib.check(lib.sgf_embedding_bag_mean(_p(tokens), tokens.stride(0), _p(ends), B, P, _p(table), _DT[table.dtype],
                                              table.stride(0), D, _p(out), _stream()), "sgf_embedding_bag_mean")
    return out


def upsample_ce_loss(logits, target, hp, wp, label_smoothing=0.0, lse_out=None, raw=False):
    """mean pixel cross-entropy of the bilinearly upsampled logits (fp32 [B,>=hp*wp,C]) against
    target int64 [B,h,w] class ids (ids outside [0,C) are ignored).  Returns (loss 0-dim, count 0-dim)."""
    lib = _lib.load()
    _req(logits, torch.float32, "logits")
    _req(target, torch.int64, "target")
    B, h, w = target.shape
    acc = torch.zeros(2, dtype=torch.float32, device=logits.device)
    args = _lib.SeglossArgs(_p(logits), logits.stride(0), logits.stride(1), B, logits.shape[2], hp, wp, h, w,
                            _p(target), float(label_smoothing), _p(acc), _p(lse_out))
    with _timed("upsample_ce", nbytes=float(target.numel() * 8)):
        _lib.check(lib.sgf_upsample_ce_loss(C.byref(args), _stream()), "sgf_upsample_ce_loss")
    if raw:
        return acc
    return acc[0] / acc[1], acc[1]


def label_propagation(features, logits, topk=3, iters=25, temperature=1.0):
    """seg_criterion.py:197-213.  features bf16 [B,P,Df] (ResNet patch features), logits fp32 [B,>=P,C] ->
    propagated class probabilities fp32 [B,P,C] and the neighbour indices int32 [B,P,topk]."""
    lib = _lib.load()
    _req(features, torch.bfloat16, "features")
    _req(logits, torch.float32, "logits")
    B, P, Df = features.shape
    Cn = logits.shape[2]
    f = features.reshape(B * P, Df)
    fn = torch.empty_like(f)
    with _timed("label_prop", nbytes=float(f.numel() * 4)):
        _lib.check(lib.sgf_l2_normalize_rows(_p(f), f.stride(0), _p(fn), fn.stride(0), B * P, Df, _stream()),
                   "sgf_l2_normalize_rows")
    Pp = (P + 3) // 4 * 4
    sim = torch.empty((B, P, Pp), dtype=torch.float32, device=f.device)
    gemm(fn, fn, sim, M=P, N=P, K=Df, batch=B, lda=Df, ldb=Df, ldc=Pp, a_batch_stride=P * Df, b_batch_stride=P * Df,
         c_batch_stride=P * Pp, tag="cosine_sim")
    nbr = torch.empty((B, P, topk), dtype=torch.int32, device=f.device)
    pa = torch.empty((B, P, Cn), dtype=torch.float32, device=f.device)
    pb = torch.empty_like(pa)
    with _timed("label_prop", nbytes=float(sim.numel() * 4 + iters * pa.numel() * 8)):
        _lib.check(lib.sgf_row_topk(_p(sim), Pp, B * P, P, topk, _p(nbr), _stream()), "sgf_row_topk")
        _lib.check(lib.sgf_label_propagation(_p(logits), logits.stride(0), logits.stride(1), B, P, Cn, float(temperature),
                                             _p(nbr), topk, iters, _p(pa), _p(pb), _stream()), "sgf_label_propagation")
    return (pa if iters % 2 == 0 else pb), nbr


# ----------------------------------------------------------------------------------------
# training path (adjoint kernels)
# ----------------------------------------------------------------------------------------
def upsample_ce_loss_bwd(logits, target, lse, count, hp, wp, dlogits, label_smoothing=0.0, grad_scale=1.0):
    """dlogits bf16 [B, tokens, ld] <- d(mean pixel CE)/d(low-res logits) * grad_scale; count = device scalar."""
    lib = _lib.load()
    _req(logits, torch.float32, "logits")
    _req(target, torch.int64, "target")
    _req(lse, torch.float32, "lse")
    _req(dlogits, torch.bfloat16, "dlogits")
    B, h, w = target.shape
    args = _lib.SeglossBwdArgs(_p(logits), logits.stride(0), logits.stride(1), B, logits.shape[2], hp, wp, h, w,
                               _p(target), _p(lse), _p(count), float(label_smoothing), float(grad_scale),
                               _p(dlogits), dlogits.stride(0), dlogits.stride(1), dlogits.shape[1])
    with _timed("upsample_ce_bwd", nbytes=float(target.numel() * 12)):
        _lib.check(lib.sgf_upsample_ce_loss_bwd(C.byref(args), _stream()), "sgf_upsample_ce_loss_bwd")
    return dlogits


def row_layernorm_bwd(*, rows, D, x=None, ldx=None, gather_idx=None, x_act=ACT_NONE, pre_add=None, g1=None, v=None,
                      g2=None, dy2=None, dv_in=None, d_res=None, dx=None, dx_accumulate=False, dg1=None, db1=None,
                      dg2=None, db2=None, d_pre_add=None, seg=None, dx_colsum=None, drop=None):
    """Adjoint of row_layernorm (see sgf_row_layernorm_bwd in include/segofa_b200.h)."""
    lib = _lib.load()
    seg_len, seg_stride, seg_off = seg if seg is not None else (0, 0, 0)

    def ld(t):
        return t.stride(-2) if t is not None else 0

    def dt(t):
        return _DT[t.dtype] if t is not None else SGF_BF16

    args = _lib.RowLnBwdArgs(
        _p(x), ld(x) if ldx is None else ldx, dt(x), _p(gather_idx), int(x_act), _p(pre_add), _p(g1),
        _p(v), ld(v), dt(v), _p(g2), _p(dy2), ld(dy2), dt(dy2), _p(dv_in), ld(dv_in), _p(d_res), ld(d_res),
        _p(dx), ld(dx), dt(dx), 1 if dx_accumulate else 0, _p(dg1), _p(db1), _p(dg2), _p(db2), _p(d_pre_add),
        rows, D, seg_len, seg_stride, seg_off, _p(dx_colsum), *_drop_fields(drop))
    nb = rows * D * sum(t.element_size() for t in (x, v, dy2, dv_in, d_res, dx) if t is not None)
    with _timed("row_layernorm_bwd" + (f":D{D}{'g1' if g1 is not None else ''}" if _TIMER is not None and _TIMER.fine else ""), nbytes=float(nb)):
        _lib.check(lib.sgf_row_layernorm_bwd(C.byref(args), _stream()), "sgf_row_layernorm_bwd")


def transpose_cast(x, M=None, N=None, out_t=None, out_c=None, colsum=None, want_t=True):
    """x [M,N] fp32/bf16 (row stride x.stride(0)) -> out_t bf16 [N, pad8(M)] (and/or out_c bf16 [M,N], colsum fp32 [N] +=)."""
    lib = _lib.load()
    _req(x, None, "x")
    M = x.shape[0] if M is None else M
    N = x.shape[1] if N is None else N
    if out_t is None and want_t:
        out_t = torch.empty((N, (M + 7) // 8 * 8), dtype=torch.bfloat16, device=x.device)
    with _timed("transpose_cast", nbytes=float(M * N * (x.element_size() + 2))):
        _lib.check(lib.sgf_transpose_cast(_p(x), _DT[x.dtype], x.stride(0), M, N, _p(out_t),
                                          out_t.stride(0) if out_t is not None else 0, _p(out_c),
                                          out_c.stride(0) if out_c is not None else 0, _p(colsum), _stream()),
                   "sgf_transpose_cast")
    return out_t


def attention_bwd(q, k, v, out, dout, dq, dk, dv, *, B, H, Tq, Tk, q_strides, k_strides, v_strides, o_strides,
                  do_strides, dq_strides, dk_strides, dv_strides, lse, delta, bias=None, head_scale=None,
                  d_head_scale=None, key_padding_mask=None, causal=False, dq_scale=1.0, dbias=None, bias_t=None):
    """bias: the fp16 tensor the forward streamed; bias_t: optional key-major copy (transpose_bias) for the dK/dV kernel."""
    lib = _lib.load()
    for t, n in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (dout, "dout"), (dq, "dq"), (dk, "dk"), (dv, "dv")):
        _req(t, torch.bfloat16, n)
    _req(lse, torch.float32, "lse")
    _req(delta, torch.float32, "delta")
    if bias is not None:
        _req(bias, torch.float16, "bias")  # the tensor the forward kernel streamed
        if bias_t is None:
            bias_t = transpose_bias(bias, Tk)
    args = _lib.AttentionBwdArgs(
        _p(q), q_strides[0], q_strides[1], _p(k), k_strides[0], k_strides[1], _p(v), v_strides[0], v_strides[1],
        _p(out), o_strides[0], o_strides[1], _p(dout), do_strides[0], do_strides[1],
        _p(dq), dq_strides[0], dq_strides[1], _p(dk), dk_strides[0], dk_strides[1], _p(dv), dv_strides[0], dv_strides[1],
        _p(bias), bias.stride(0) if bias is not None else 0, bias.stride(1) if bias is not None else 0,
        _p(head_scale), _p(d_head_scale), _p(key_padding_mask), _p(lse), _p(delta), float(dq_scale),
        B, H, Tq, Tk, 1 if causal else 0, _p(dbias),
        _p(bias_t), bias_t.stride(0) if bias_t is not None else 0, bias_t.stride(1) if bias_t is not None else 0)
    if bias_t is not None:
        _req(bias_t, torch.float16, "bias_t")
    pairs = Tq * Tk if not causal else Tq * (Tq + 1) / 2.0
    with _timed("attention_bwd_tcgen05", 10.0 * B * H * pairs * 64):  # 5 algorithmic tile GEMMs
        _lib.check(lib.sgf_attention_bwd_bf16(C.byref(args), _stream()), "sgf_attention_bwd_bf16")


def transpose_bias(bias, Tk):
    """key-major copy [H, Tk_pad, roundup(Tq, 64)] of an fp16 bias [H, Tq, row_stride] (for attention_bwd's bias_t)."""
    lib = _lib.load()
    _req(bias, torch.float16, "bias")
    H, Tq, ld = bias.shape
    assert bias.stride(2) == 1 and ld % 8 == 0
    cols = (Tk + 7) // 8 * 8
    out = torch.empty((H, cols, (Tq + 63) // 64 * 64), dtype=torch.float16, device=bias.device)
    with _timed("attn_bias", nbytes=4.0 * H * Tq * cols):
        _lib.check(lib.sgf_transpose16_batched(_p(bias), bias.stride(0), bias.stride(1), H, Tq, cols, _p(out),
                                               out.stride(0), out.stride(1), _stream()), "sgf_transpose16_batched")
    return out


def bias_block_csr(bucket, ids, lo, row_stride):
    """Static CSR grouping of the positions of one relative-position block by bucket (host-side, once per shape):
    returns (order int32 [n*n] of flat positions i*row_stride + j, offsets int32 [num_rel+1] -- sized by the caller's
    table) as device tensors.  bucket int64 [*,*], ids int64 [n]."""
    b = bucket[ids][:, ids]  # [n, n] bucket of (i, j)
    n = ids.numel()
    flat = b.reshape(-1)
    perm = torch.argsort(flat, stable=True)
    ii, jj = perm // n, perm % n
    order = ((ii + lo) * row_stride + (jj + lo)).to(torch.int32)
    return order.contiguous(), flat


def attn_bias_bwd(dbias, blocks=(), dabs_acc=None):
    """Adjoint of build_attn_bias for one layer (see sgf_attn_bias_bwd): dbias fp32 [H,Tq,row_stride] is consumed and
    cleared; blocks: iterable of (order int32, offsets int32 [num_rel+1], dtable fp32 [num_rel,H])."""
    lib = _lib.load()
    _req(dbias, torch.float32, "dbias")
    H, Tq, _ = dbias.shape
    args = _lib.BiasBwdArgs()
    args.dbias = dbias.data_ptr()
    args.dabs_acc = dabs_acc.data_ptr() if dabs_acc is not None else None
    args.head_stride, args.row_stride = dbias.stride(0), dbias.stride(1)
    args.H, args.Tq, args.num_blocks = H, Tq, len(blocks)
    for i, (order, offsets, dtable) in enumerate(blocks):
        _req(dtable, torch.float32, "dtable")
        _req(order, torch.int32, "order")
        _req(offsets, torch.int32, "offsets")
        assert dtable.is_contiguous() and dtable.shape[1] == H and offsets.numel() == dtable.shape[0] + 1
        args.order[i], args.offsets[i], args.dtable[i] = order.data_ptr(), offsets.data_ptr(), dtable.data_ptr()
        args.num_rel[i] = dtable.shape[0]
    with _timed("attn_bias_bwd", nbytes=20.0 * H * Tq * dbias.stride(1)):
        _lib.check(lib.sgf_attn_bias_bwd(C.byref(args), _stream()), "sgf_attn_bias_bwd")


def artificial_sample(name_tokens, name_lens, B, hp, S, *, lo=1, hi=33, seed=1, step=None, seg_id_offset=59457, eos_id=2,
                      pad_id=1, want_grid=False):
    """Image-free training sample generated on the device (segmentation_dataset.py:303-329, 'rand_k-lo-hi').
    name_tokens int64 [C, Lmax] / name_lens int32 [C]: BPE ids of the class names.  Returns (bag_tokens int64
    [B, hp*hp*Lmax], bag_ends int64 [B*hp*hp], text2seg_target int64 [B, S*S+1]) (+ the label grids when want_grid)."""
    lib = _lib.load()
    _req(name_tokens, torch.int64, "name_tokens")
    _req(name_lens, torch.int32, "name_lens")
    Cn, Lmax = name_tokens.shape
    P = hp * hp
    dev = name_tokens.device
    bag = torch.empty((B, P * Lmax), dtype=torch.int64, device=dev)
    ends = torch.empty((B * P,), dtype=torch.int64, device=dev)
    target = torch.empty((B, S * S + 1), dtype=torch.int64, device=dev)
    grid = torch.zeros((B, 1026), dtype=torch.int32, device=dev) if want_grid else None
    args = _lib.ArtSampleArgs(_p(name_tokens), name_tokens.stride(0), _p(name_lens), Cn, B, hp, S, lo, hi, int(seed) & 0xFFFFFFFF,
                              _p(step), seg_id_offset, eos_id, pad_id, _p(bag), bag.stride(0), _p(ends), _p(target), _p(grid))
    with _timed("artificial_sample", nbytes=float(target.numel() * 8 + bag.numel() * 8)):
        _lib.check(lib.sgf_artificial_sample(C.byref(args), _stream()), "sgf_artificial_sample")
    return (bag, ends, target, grid) if want_grid else (bag, ends, target)


def adam_step(param, grad, exp_avg, exp_avg_sq, *, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, step=1,
              step_dev=None, grad_scale=None):
    lib = _lib.load()
    for t, n in ((param, "param"), (grad, "grad"), (exp_avg, "exp_avg"), (exp_avg_sq, "exp_avg_sq")):
        _req(t, torch.float32, n)
    with _timed("adam", nbytes=float(param.numel() * 28)):
        _lib.check(lib.sgf_adam_step(_p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), param.numel(), float(lr),
                                     float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
                                     _p(step_dev), _p(grad_scale), _stream()), "sgf_adam_step")


def sumsq(x, out):
    lib = _lib.load()
    _req(x, torch.float32, "x")
    with _timed("sumsq", nbytes=float(x.numel() * 4)):
        _lib.check(lib.sgf_sumsq(_p(x), x.numel(), _p(out), _stream()), "sgf_sumsq")
    return out


def image_prep_u8(image_u8, rs_hw, *, crop=None, flip=False, mean, std, out=None):
    """uint8 [H, W, 3] on the device -> fp32 [3, out_h, out_w]: cv2 INTER_LINEAR resize to rs_hw, window `crop` =
    (y, x, h, w) of it, horizontal flip, /255, (x - mean) / std -- bit-identical to cv2 + torchvision
    (data/mm_data/segmentation_dataset.py:155-156, 236-240, 253-257)."""
    lib = _lib.load()
    _req(image_u8, torch.uint8, "image_u8")
    assert image_u8.dim() == 3 and image_u8.shape[2] == 3 and image_u8.stride(2) == 1 and image_u8.stride(1) == 3
    H, W = image_u8.shape[:2]
    rs_h, rs_w = rs_hw
    cy, cx, oh, ow = crop if crop is not None else (0, 0, rs_h, rs_w)
    if out is None:
        out = torch.empty((3, oh, ow), dtype=torch.float32, device=image_u8.device)
    _req(out, torch.float32, "out")
    assert tuple(out.shape) == (3, oh, ow) and out.stride(2) == 1
    args = _lib.ImagePrepArgs(_p(image_u8), image_u8.stride(0), H, W, rs_h, rs_w, cy, cx, oh, ow, 1 if flip else 0,
                              (C.c_float * 3)(*mean), (C.c_float * 3)(*std), _p(out), out.stride(0), out.stride(1))
    with _timed("image_prep", nbytes=float(image_u8.numel() + out.numel() * 4)):
        _lib.check(lib.sgf_image_prep_u8(C.byref(args), _stream()), "sgf_image_prep_u8")
    return out


def segmap_prep_u8(seg_u8, num_seg, rs_hw, grid_hw, *, crop=None, flip=False, seg_id_offset=59457, bos_id=0, eos_id=2,
                   want_downsampled=False, want_ori=False):
    """raw uint8 label map [H, W] on the device -> (target int64 [out_h*out_w+1], prev_output_tokens int64 [grid+1],
    downsampled_target or None, ori_classes int64 [H, W] or None)  (segmentation_dataset.py:224-227, 241-265)."""
    lib = _lib.load()
    _req(seg_u8, torch.uint8, "seg_u8")
    assert seg_u8.dim() == 2 and seg_u8.stride(1) == 1
    H, W = seg_u8.shape
    rs_h, rs_w = rs_hw
    cy, cx, oh, ow = crop if crop is not None else (0, 0, rs_h, rs_w)
    gh, gw = grid_hw
    dev = seg_u8.device
    target = torch.empty((oh * ow + 1,), dtype=torch.int64, device=dev)
    prev = torch.empty((gh * gw + 1,), dtype=torch.int64, device=dev)
    down = torch.empty((gh * gw + 1,), dtype=torch.int64, device=dev) if want_downsampled else None
    ori = torch.empty((H, W), dtype=torch.int64, device=dev) if want_ori else None
    args = _lib.SegmapPrepArgs(_p(seg_u8), seg_u8.stride(0), H, W, num_seg, rs_h, rs_w, cy, cx, oh, ow, 1 if flip else 0,
                               gh, gw, seg_id_offset, bos_id, eos_id, _p(target), _p(prev), _p(down), _p(ori))
    with _timed("segmap_prep", nbytes=float(seg_u8.numel() + 8 * (target.numel() + prev.numel()))):
        _lib.check(lib.sgf_segmap_prep_u8(C.byref(args), _stream()), "sgf_segmap_prep_u8")
    return target, prev, down, ori
