"""Flat execution plan of the segofa forward on one B200: a sequence of C-ABI kernel launches
(ifseg_b200.ops -> libsegofa_b200.so) over device buffers owned by PyTorch's caching allocator.

Data layout in HBM
  * activations are batch-major row matrices [B*T, D]; the residual stream is fp32, every GEMM
    operand (LayerNorm output, attention output, FFN hidden) is bf16;
  * q/k/v live in one [B*T, 3D] bf16 buffer written by the fused QKV GEMM (q pre-scaled by
    (2 d_h)^-1/2 in its epilogue) and are read in place by the attention kernel through 4-D TMA
    maps (d, head, token, batch) -- no head transposes are materialised;
  * the stem runs NHWC bf16; FrozenBatchNorm is folded to a per-channel (scale, bias) applied in
    the convolution epilogue together with ReLU and the residual add;
  * the additive attention bias is batch-invariant and parameter-only: it is built once per
    forward as fp16 [H, Tq, Tk_pad] per layer (abs-pos q_pos k_pos^T by a head-batched GEMM +
    rel-pos table lookups), never per batch element and never through clone/cat/interpolate
    chains (those are 55-60% of the reference's forward time).

Reference semantics followed (file:line under /root/reference/models/segofa/):
  encoder_module.py:677-851 (encode), :499-675 (encode_with_artificial_image),
  decoder_module.py:486-677 (surrogate decoder), :290-294 (seg_projection),
  unify_transformer_layer.py:222-292, 431-581, unify_multihead_attention.py:327-523,
  resnet.py:215-229, frozen_bn.py:40-45.
"""
import os
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import ops
from .config import SegOFAConfig

_BF16 = torch.bfloat16


def _pad64(n):
    return (n + 63) // 64 * 64


class _ConvW:
    __slots__ = ("w", "scale", "bias", "kh", "kw", "stride", "pad", "cin", "cout", "k", "ldk")


class SegOFAEngine:
    def __init__(self, model, live=None):
        # live = a SegOFATrainEngine: trainable operands are then VIEWS of its bf16 operand buffers / fp32 master
        # arena, so this engine follows every optimizer step without being rebuilt (the trainer's no-grad
        # real-image pass, seg_criterion.py:185).  ffn_layernorm is not folded in that mode (derived weights).
        self.live = live
        self.cfg: SegOFAConfig = model.cfg
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError(
                "segofa_b200: the model is on %s; the hot path is CUDA-only (sm_100a) and has no CPU fallback -- "
                "move the model to a B200 with .cuda()" % p.device
            )
        from . import _lib

        _lib.load()  # fail loudly if the extension is missing
        self.device = p.device
        self.model = model
        self._shape_cache: Dict = {}
        # The additive position bias is a pure function of the (frozen) parameters and the token grid,
        # exactly like the folded BN / fused QKV weights prepared below: it is derived once per shape and
        # dropped with the engine whenever the model's parameters may change (SegOFAModel.invalidate_engine).
        self.cache_position_bias = True
        # ffn_layernorm folded into the fc1/fc2 epilogues (default) or run as a separate row kernel
        self.fold_ffn_layernorm = os.environ.get("SGF_FOLD_FFN_LN", "1") != "0" and live is None
        self._bias_cache: Dict = {}
        with torch.no_grad():
            self._prepare(model)

    # ------------------------------------------------------------------------------------
    # parameter preparation (derived device tensors; rebuilt whenever the model invalidates)
    # ------------------------------------------------------------------------------------
    def _f32(self, t):
        if self.live is not None and self.live.arena.has(t):
            return self.live.arena.value(t)  # live view of the fp32 master (== t.data only for an fp32 model)
        return t.detach().to(device=self.device, dtype=torch.float32).contiguous()

    def _b16(self, t):
        if self.live is not None:
            d = self.live.dense_index.get((id(t),))
            if d is not None:
                return d.w16
        return t.detach().to(device=self.device, dtype=_BF16).contiguous()

    def _fused(self, weights, biases):
        """bf16 [sum N, K] weight and fp32 bias of several nn.Linear fused along the output dimension."""
        if self.live is not None:
            d = self.live.dense_index.get(tuple(id(w) for w in weights))
            if d is not None:
                return d.w16, d.b32
        return (self._b16(torch.cat([w.detach() for w in weights], 0)),
                self._f32(torch.cat([b.detach() for b in biases], 0)))

    def _ln(self, m):
        return (self._f32(m.weight), self._f32(m.bias))

    def _conv(self, conv, bn) -> _ConvW:
        c = _ConvW()
        w = conv.weight.detach().float()
        c.cout, c.cin, c.kh, c.kw = w.shape
        c.stride, c.pad = conv.stride[0], conv.padding[0]
        if c.cin % 8 == 0:
            # [Cout, ky, kx, cin] tap-major: the K order of the implicit-GEMM convolutions and of a 1x1 GEMM alike
            c.k = c.kh * c.kw * c.cin
            wm = w.permute(0, 2, 3, 1).reshape(c.cout, c.k)
        else:
            # conv1 (7x7/2 over 3 channels): 7 filter-row taps of an 8-pixel x 8-channel window, zero weights for the
            # eighth pixel and for channels 3..7 (ops.conv2d window mode) -- K = 7 * 64
            assert c.kw <= 8 and c.cin <= 8
            c.k = c.kh * 64
            wm = torch.zeros(c.cout, c.kh, 8, 8, dtype=torch.float32, device=w.device)
            wm[:, :, : c.kw, : c.cin] = w.permute(0, 2, 3, 1)
            wm = wm.reshape(c.cout, c.k)
        c.ldk = c.k
        c.w = self._b16(wm.contiguous())
        scale = bn.weight.detach().float() * (bn.running_var.detach().float() + bn.eps).rsqrt()  # frozen_bn.py:40-41
        c.scale = self._f32(scale)
        c.bias = self._f32(bn.bias.detach().float() - bn.running_mean.detach().float() * scale)
        return c

    def _attn(self, m, cross=False):
        d = {}
        if cross:
            d["wq"], d["bq"] = self._b16(m.q_proj.weight), self._f32(m.q_proj.bias)
            if self.live is None:  # (live: only the all-layer fused operand w_cross_kv_all exists)
                d["wkv"], d["bkv"] = self._fused([m.k_proj.weight, m.v_proj.weight], [m.k_proj.bias, m.v_proj.bias])
        else:
            d["wqkv"], d["bqkv"] = self._fused([m.q_proj.weight, m.k_proj.weight, m.v_proj.weight],
                                               [m.q_proj.bias, m.k_proj.bias, m.v_proj.bias])
        d["wo"], d["bo"] = self._b16(m.out_proj.weight), self._f32(m.out_proj.bias)
        d["c_attn"] = self._f32(m.c_attn) if m.c_attn is not None else None
        return d

    def _ffn_fold(self, layer):
        """ffn_layernorm folded into fc2 (see sgf_gemm_args row-norm): W2' = W2 * gamma, u = rowsum(W2'),
        c = b2 + W2 beta."""
        if self.live is not None:
            return dict(w2=self._b16(layer.fc2.weight), b2=self._f32(layer.fc2.bias), ln_ffn=self._ln(layer.ffn_layernorm))
        w2 = layer.fc2.weight.detach().float().to(self.device)
        g, b = layer.ffn_layernorm.weight.detach().float().to(self.device), layer.ffn_layernorm.bias.detach().float().to(self.device)
        w2f = (w2 * g.unsqueeze(0)).to(_BF16).contiguous()
        u = w2f.float().sum(dim=1).contiguous()
        c = (layer.fc2.bias.detach().float().to(self.device) + w2 @ b).contiguous()
        return dict(w2f=w2f, u2=u, c2=c, w2=self._b16(layer.fc2.weight), b2=self._f32(layer.fc2.bias),
                    ln_ffn=self._ln(layer.ffn_layernorm))

    def _prepare(self, model):
        enc, dec, cfg = model.encoder, model.decoder, self.cfg
        s = enc.embed_images
        self.stem_conv1 = self._conv(s.conv1, s.bn1)
        self.stem_blocks = []
        for li in (1, 2, 3):
            for blk in getattr(s, f"layer{li}"):
                b = dict(c1=self._conv(blk.conv1, blk.bn1), c2=self._conv(blk.conv2, blk.bn2),
                         c3=self._conv(blk.conv3, blk.bn3), down=None, stride=blk.stride)
                if blk.downsample is not None:
                    b["down"] = self._conv(blk.downsample[0], blk.downsample[1])
                self.stem_blocks.append(b)
        type_emb = self._f32(enc.type_embedding.weight)
        self.type_txt, self.type_img = type_emb[0].contiguous(), type_emb[1].contiguous()
        self.embed_tokens = enc.embed_tokens.weight.detach()
        if self.embed_tokens.dtype not in (torch.float32, _BF16) or not self.embed_tokens.is_contiguous():
            self.embed_tokens = self._f32(self.embed_tokens)
        self.w_image_proj, self.b_image_proj = self._b16(enc.image_proj.weight), self._f32(enc.image_proj.bias)
        self.ln_emb, self.ln_patch = self._ln(enc.layernorm_embedding), self._ln(enc.patch_layernorm_embedding)
        self.enc_pos_table = self._f32(enc.embed_positions.weight)
        self.enc_img_pos_table = self._f32(enc.embed_image_positions.weight)
        self.ln_pos, self.ln_img_pos = self._ln(enc.pos_ln), self._ln(enc.image_pos_ln)
        self.w_pos_q, self.b_pos_q = self._b16(enc.pos_q_linear.weight), self._f32(enc.pos_q_linear.bias)
        self.w_pos_k, self.b_pos_k = self._b16(enc.pos_k_linear.weight), self._f32(enc.pos_k_linear.bias)
        self.enc_layers = []
        for l in enc.layers:
            self.enc_layers.append(dict(
                attn=self._attn(l.self_attn), ln_self=self._ln(l.self_attn_layer_norm), ln_attn=self._ln(l.attn_ln),
                ln_final=self._ln(l.final_layer_norm),
                w1=self._b16(l.fc1.weight), b1=self._f32(l.fc1.bias), **self._ffn_fold(l)))
        self.ln_enc_out = self._ln(enc.layer_norm)
        self.enc_tok_rel = [self._f32(t.weight) for t in enc.token_rel_pos_table_list]
        self.enc_img_rel = [self._f32(t.weight) for t in enc.image_rel_pos_table_list]
        self.token_rp_bucket = enc.token_rp_bucket.to(self.device)
        self.image_rp_bucket = enc.image_rp_bucket.to(self.device)
        # decoder
        self.dec_ln_emb = self._ln(dec.layernorm_embedding)
        self.seg_pos_table = self._f32(dec.embed_seg_positions.weight)
        self.ln_seg_pos = self._ln(dec.seg_pos_ln)
        self.w_self_pq, self.b_self_pq = self._b16(dec.self_pos_q_linear.weight), self._f32(dec.self_pos_q_linear.bias)
        self.w_self_pk, self.b_self_pk = self._b16(dec.self_pos_k_linear.weight), self._f32(dec.self_pos_k_linear.bias)
        self.w_cross_pq, self.b_cross_pq = self._b16(dec.cross_pos_q_linear.weight), self._f32(dec.cross_pos_q_linear.bias)
        self.w_cross_pk, self.b_cross_pk = self._b16(dec.cross_pos_k_linear.weight), self._f32(dec.cross_pos_k_linear.bias)
        self.dec_layers = []
        for l in dec.layers:
            self.dec_layers.append(dict(
                attn=self._attn(l.self_attn), cross=self._attn(l.encoder_attn, cross=True),
                ln_self=self._ln(l.self_attn_layer_norm), ln_self_attn=self._ln(l.self_attn_ln),
                ln_enc_attn=self._ln(l.encoder_attn_layer_norm), ln_cross_attn=self._ln(l.cross_attn_ln),
                ln_final=self._ln(l.final_layer_norm),
                w1=self._b16(l.fc1.weight), b1=self._f32(l.fc1.bias), **self._ffn_fold(l)))
        # all decoder layers' cross-attention K/V projections of encoder_out as ONE GEMM (N = L*2D)
        if self.live is not None:
            self.w_cross_kv_all, self.b_cross_kv_all = self._fused(
                [w for l in dec.layers for w in (l.encoder_attn.k_proj.weight, l.encoder_attn.v_proj.weight)],
                [b for l in dec.layers for b in (l.encoder_attn.k_proj.bias, l.encoder_attn.v_proj.bias)])
        else:
            self.w_cross_kv_all = torch.cat([d["cross"]["wkv"] for d in self.dec_layers], 0).contiguous()
            self.b_cross_kv_all = torch.cat([d["cross"]["bkv"] for d in self.dec_layers], 0).contiguous()
        self.ln_dec_out = self._ln(dec.layer_norm)
        self.dec_seg_rel = [self._f32(t.weight) for t in dec.seg_rel_pos_table_list]
        self.seg_rp_bucket = dec.seg_rp_bucket.to(self.device)
        self.w_seg_proj = self._b16(dec.seg_projection.weight)
        self.seg_bucket_size = dec.seg_bucket_size

    # ------------------------------------------------------------------------------------
    # stem: resnet.py:215-229
    # ------------------------------------------------------------------------------------
    def _conv_gemm(self, x, c: _ConvW, act, residual=None):
        """x [N,H,W,Cin] bf16 NHWC -> [N,Ho,Wo,Cout] bf16.  Every convolution is an im2col-free implicit GEMM: 1x1/1 is a
        plain GEMM over the pixel rows, everything else (3x3/1, 3x3/2, 1x1/2) goes through the TMA-box convolution."""
        n, h, w, _ = x.shape
        if c.kh == 1 and c.stride == 1:
            out = torch.empty((n, h, w, c.cout), dtype=_BF16, device=x.device)
            ops.gemm(x.view(n * h * w, c.cin), c.w, out.view(-1, c.cout), M=n * h * w, N=c.cout, K=c.k, lda=c.cin, ldb=c.ldk,
                     scale=c.scale, bias=c.bias, act=act,
                     residual=residual.view(-1, c.cout) if residual is not None else None,
                     tag=f"stem1x1s1_{h}_k{c.k}_n{c.cout}")
            return out
        assert residual is None
        if c.kh == 3 and c.stride == 1:
            return ops.conv3x3_s1(x, c.w, c.scale, c.bias, act=act, tag=f"stem3x3_{h}")
        return ops.conv2d(x, c.w, c.scale, c.bias, kh=c.kh, kw=c.kw, stride=c.stride, pad=c.pad, act=act,
                          tag=f"stem{c.kh}x{c.kw}s{c.stride}_{h}")

    def stem(self, patch_images):
        c1 = self.stem_conv1
        _, _, H, W = patch_images.shape
        # conv1 7x7/2 (resnet.py:189-213): the image goes to a zero-padded 8-channel NHWC buffer once; the convolution reads
        # it as overlapping 128-byte windows (8 pixels x 8 channels) -- 7 k-blocks, no patch matrix
        hp, wp = H + 2 * c1.pad, (W + 2 * c1.pad + 8 + 7) // 8 * 8
        xp = ops.nchw_to_nhwc8_padded(patch_images.float(), c1.pad, hp, wp)
        x = ops.conv2d(xp, c1.w, c1.scale, c1.bias, kh=c1.kh, kw=1, stride=c1.stride, act=ops.ACT_RELU,
                       window=(H, W, c1.pad, c1.cin, c1.kw), tag="stem_conv1")
        x = ops.maxpool3x3s2(x)
        for b in self.stem_blocks:
            idn = x if b["down"] is None else self._conv_gemm(x, b["down"], ops.ACT_NONE)
            y = self._conv_gemm(x, b["c1"], ops.ACT_RELU)
            y = self._conv_gemm(y, b["c2"], ops.ACT_RELU)
            x = self._conv_gemm(y, b["c3"], ops.ACT_RELU, residual=idn)  # relu(identity + bn3(conv3)) :128-135
        return x  # [B,h,w,1024] NHWC == [B,P,1024] token-major

    # ------------------------------------------------------------------------------------
    # position bias (batch-invariant)
    # ------------------------------------------------------------------------------------
    def _cached(self, key, make):
        """Shape-dependent index tensors are built once (outside any CUDA-graph capture)."""
        t = self._shape_cache.get(key)
        if t is None:
            t = make().to(self.device)
            self._shape_cache[key] = t
        return t

    def _image_position_ids(self, h, w):
        b = self.cfg.image_bucket_size
        return self._cached(("img_ids", h, w), lambda: (
            torch.arange(w).unsqueeze(0).expand(h, w) + torch.arange(h).unsqueeze(1) * b + 1).reshape(-1))

    def _abs_bias(self, pos_q_in, pos_k_in, wq, bq, wk, bk):
        """fp32 [H, Tq, pad64(Tk)] = (pos_q W_q^T + b_q) * pos_scaling  .  (pos_k W_k^T + b_k)^T per head."""
        cfg = self.cfg
        D, H, dh = cfg.embed_dim, cfg.heads, cfg.head_dim
        Tq, Tk = pos_q_in.shape[0], pos_k_in.shape[0]
        pq = ops.gemm(pos_q_in, wq, bias=bq, alpha=cfg.pos_scaling, alpha_cols=D)
        Tkp = _pad64(Tk)
        pk = torch.zeros((Tkp, D), dtype=_BF16, device=self.device)  # zero rows beyond Tk: the GEMM below runs at width Tkp
        ops.gemm(pos_k_in, wk, pk[:Tk], bias=bk)
        out = torch.empty((H, Tq, Tkp), dtype=torch.float32, device=self.device)
        # N = padded width: key rows >= Tk are TMA zero-fill -> zero padding columns, vectorised epilogue (N % 32 == 0)
        ops.gemm(pq, pk, out, M=Tq, N=Tkp, K=dh, batch=H, lda=D, ldb=D, a_batch_stride=dh, b_batch_stride=dh,
                 ldc=Tkp, c_batch_stride=Tq * Tkp)
        return out

    def _interp_needed(self, h, w):
        oh = self.cfg.orig_patch_image_size // 16
        return (h, w) != (oh, oh)

    # General patch grids (validation keeps the image aspect ratio).  The relative-position bias of the image block is
    # defined on the orig grid and resized with two separable bilinear passes; this is a batch-invariant, parameter-only
    # precompute (cached per grid like every other bias), so it is expressed with torch on the device and handed to
    # sgf_build_attn_bias as `dense_add` -- the per-image hot path is unchanged.
    def _interp_image_rel(self, table, oh, h, w):
        """encoder_module.py:321-331, 782-808 -> fp32 [H, h*w, h*w]: key axis first, then query axis."""
        H = self.cfg.heads
        oids = self._image_position_ids(oh, oh)
        v = table[self.image_rp_bucket[oids][:, oids]].permute(2, 0, 1)            # [H, q(oh*oh), k(oh*oh)]
        v = v.reshape(H, oh * oh, oh, oh).permute(1, 0, 2, 3)                       # [(q), H, kh, kw]
        v = F.interpolate(v, size=(h, w), mode="bilinear")                          # keys -> (h, w)
        v = v.reshape(oh, oh, H, h * w).permute(3, 2, 0, 1)                         # [(k'), H, qh, qw]
        v = F.interpolate(v, size=(h, w), mode="bilinear")                          # queries -> (h, w)
        return v.reshape(h * w, H, h * w).permute(1, 2, 0)                          # [H, q', k']

    def _interp_seg_rel(self, table, sb, h, w):
        """decoder_module.py:601-625 -> fp32 [H, Td, Td]: query axis first, then key axis; bos row/column exempt."""
        H = self.cfg.heads
        n, Td = sb * sb, h * w + 1
        v = table[self.seg_rp_bucket].permute(2, 0, 1)                              # [H, q(n+1), k(n+1)]
        t = v.permute(2, 0, 1)                                                      # [k, H, q]
        seg = F.interpolate(t[..., 1:].reshape(n + 1, H, sb, sb), size=(h, w), mode="bilinear").reshape(n + 1, H, h * w)
        t = torch.cat([t[..., :1], seg], dim=-1).permute(2, 1, 0)                   # [q'(Td), H, k(n+1)]
        seg = F.interpolate(t[..., 1:].reshape(Td, H, sb, sb), size=(h, w), mode="bilinear").reshape(Td, H, h * w)
        return torch.cat([t[..., :1], seg], dim=-1).permute(1, 0, 2)               # [H, q', k'(Td)]

    def _encoder_bias(self, h, w, T_txt, artificial):
        """list of fp32 [H,T_e,pad64(T_e)] per layer + the post-LN position embeddings [T_e,D] bf16."""
        key = ("enc", h, w, T_txt, artificial)
        if self.cache_position_bias and key in self._bias_cache:
            return self._bias_cache[key]
        cfg = self.cfg
        P, D = h * w, cfg.embed_dim
        T = P + T_txt
        oh = cfg.orig_patch_image_size // 16
        if artificial and P > oh * oh:
            raise NotImplementedError("segofa_b200: the image-free branch runs on the patch_image_size grid only")
        interp = (not artificial) and self._interp_needed(h, w)  # validation keeps the aspect ratio: general grids
        pos = torch.empty((T, D), dtype=_BF16, device=self.device)
        if P > oh * oh:
            # encoder_module.py:358-370: the orig-grid absolute position embeddings, bilinearly resized to (h, w)
            oids = self._image_position_ids(oh, oh)
            old = self.enc_img_pos_table[oids].reshape(1, oh, oh, D).permute(0, 3, 1, 2)
            new = F.interpolate(old, size=(h, w), mode="bilinear").permute(0, 2, 3, 1).reshape(P, D).contiguous()
            ops.row_layernorm(new, rows=P, ln2=self.ln_img_pos, out2=pos)
        else:
            ops.row_layernorm(self.enc_img_pos_table, rows=P, gather_idx=self._image_position_ids(h, w),
                              ln2=self.ln_img_pos, out2=pos)
        ops.row_layernorm(self.enc_pos_table, rows=T_txt, ln2=self.ln_pos, out2=pos[P:])
        absb = self._abs_bias(pos, pos, self.w_pos_q, self.b_pos_q, self.w_pos_k, self.b_pos_k)
        tok_ids = self._cached(("arange", T_txt), lambda: torch.arange(T_txt))
        ids = self._image_position_ids(h, w) if not interp else None
        biases = []
        for l in range(cfg.enc_layers):
            tok_block = (self.token_rp_bucket, tok_ids, self.enc_tok_rel[l], P, T)
            if not interp:
                blocks = [(self.image_rp_bucket, ids, self.enc_img_rel[l], 0, P), tok_block]
                biases.append(ops.build_attn_bias(absb, T, blocks, f16=True))  # fp16: what the attention kernel streams
            else:
                dense = torch.zeros_like(absb)
                dense[:, :P, :P] = self._interp_image_rel(self.enc_img_rel[l], oh, h, w)
                biases.append(ops.build_attn_bias(absb, T, [tok_block], dense_add=dense, f16=True))
        res = (biases, pos)
        if self.cache_position_bias:
            self._bias_cache[key] = res
        return res

    def _decoder_bias(self, h, w, enc_pos):
        key = ("dec", h, w, enc_pos.shape[0])
        if self.cache_position_bias and key in self._bias_cache:
            return self._bias_cache[key]
        cfg = self.cfg
        sb = self.seg_bucket_size
        interp = (h, w) != (sb, sb)
        Td, D = h * w + 1, cfg.embed_dim
        tgt_pos = torch.empty((Td, D), dtype=_BF16, device=self.device)
        if not interp:
            ops.row_layernorm(self.seg_pos_table, rows=Td, ln2=self.ln_seg_pos, out2=tgt_pos)  # ids 0..n == table rows
        else:
            # decoder_module.py:541-550: the seg_bucket-grid embeddings resized to (h, w); slot 0 (bos) carried through
            old = self.seg_pos_table[1: sb * sb + 1].reshape(1, sb, sb, D).permute(0, 3, 1, 2)
            new = F.interpolate(old, size=(h, w), mode="bilinear").permute(0, 2, 3, 1).reshape(h * w, D)
            ops.row_layernorm(torch.cat([self.seg_pos_table[0:1], new], 0).contiguous(), rows=Td, ln2=self.ln_seg_pos,
                              out2=tgt_pos)
        self_abs = self._abs_bias(tgt_pos, tgt_pos, self.w_self_pq, self.b_self_pq, self.w_self_pk, self.b_self_pk)
        cross_abs = self._abs_bias(tgt_pos, enc_pos, self.w_cross_pq, self.b_cross_pq, self.w_cross_pk, self.b_cross_pk)
        seg_ids = self._cached(("arange", Td), lambda: torch.arange(Td))
        if not interp:
            self_biases = [ops.build_attn_bias(self_abs, Td, [(self.seg_rp_bucket, seg_ids, self.dec_seg_rel[l], 0, Td)],
                                               f16=True) for l in range(cfg.dec_layers)]
        else:
            self_biases = []
            for l in range(cfg.dec_layers):
                dense = torch.zeros_like(self_abs)
                dense[:, :, :Td] = self._interp_seg_rel(self.dec_seg_rel[l], sb, h, w)
                self_biases.append(ops.build_attn_bias(self_abs, Td, (), dense_add=dense, f16=True))
        cross_abs = ops.build_attn_bias(cross_abs, enc_pos.shape[0], (), f16=True)
        res = (self_biases, cross_abs)
        if self.cache_position_bias:
            self._bias_cache[key] = res
        return res

    # ------------------------------------------------------------------------------------
    # transformer blocks
    # ------------------------------------------------------------------------------------
    def _self_attention(self, a, L, B, T, bias, causal, kpm):
        cfg = self.cfg
        D, H = cfg.embed_dim, cfg.heads
        qkv = ops.gemm(a, L["wqkv"], bias=L["bqkv"], alpha=cfg.attn_scaling, alpha_cols=D, tag="qkv")  # [B*T,3D]
        o = torch.empty((B * T, D), dtype=_BF16, device=self.device)
        ops.attention(qkv, qkv[:, D:], qkv[:, 2 * D:], o, B=B, H=H, Tq=T, Tk=T, q_strides=(3 * D, T * 3 * D),
                      k_strides=(3 * D, T * 3 * D), v_strides=(3 * D, T * 3 * D), o_strides=(D, T * D), bias=bias,
                      head_scale=L["c_attn"], key_padding_mask=kpm, causal=causal)
        return ops.gemm(o, L["wo"], bias=L["bo"], out_dtype=torch.float32, tag="out_proj")

    def _ffn(self, a, x, L, stats):
        """x <- x + fc2(ffn_layernorm(gelu(fc1(a)))) with the F-wide LayerNorm folded into the two GEMM
        epilogues: fc1 writes per-row (sum, sumsq) of each 64-column block of its bf16 output into `stats`,
        fc2 sums them in order and applies rstd * (acc - mean * u) + c."""
        if not self.fold_ffn_layernorm:
            f = ops.gemm(a, L["w1"], bias=L["b1"], act=ops.ACT_GELU, tag="fc1")
            g = torch.empty_like(f)
            ops.row_layernorm(f, ln2=L["ln_ffn"], out2=g)
            ops.gemm(g, L["w2"], x, bias=L["b2"], residual=x, tag="fc2")
            return
        f = ops.gemm(a, L["w1"], bias=L["b1"], act=ops.ACT_GELU, rowstats_out=stats, tag="fc1")
        ops.gemm(f, L["w2f"], x, bias=L["c2"], residual=x, rownorm=(stats, L["u2"], self.cfg.ffn_dim), tag="fc2")

    # ------------------------------------------------------------------------------------
    # encoder
    # ------------------------------------------------------------------------------------
    def encode(self, src_tokens, patch_images=None, patch_masks=None, bag_tokens=None, bag_offsets=None,
               has_pads: Optional[bool] = None):
        cfg = self.cfg
        dev = self.device
        D = cfg.embed_dim
        src_tokens = src_tokens.to(dev)
        B, T_txt = src_tokens.shape
        artificial = bag_tokens is not None
        feat = None
        bag = None
        if artificial:
            # image-free branch (encoder_module.py:529-551): every patch is the mean token embedding of its
            # category word; no ResNet, no image_proj
            h = w = cfg.patch_image_size // 16
            bag = ops.embedding_bag_mean(bag_tokens.to(dev).contiguous(), bag_offsets.to(dev).contiguous(),
                                         self.embed_tokens, h * w)
        elif patch_images is None:
            raise NotImplementedError("segofa_b200: text-only encoding is not on the IFSeg hot path")
        else:
            feat = self.stem(patch_images.to(dev))
            h, w = feat.shape[1], feat.shape[2]
        P = h * w
        T = P + T_txt
        # padding: encoder_module.py:730,737-742
        pad = torch.zeros((B, T), dtype=torch.uint8, device=dev)
        pad[:, P:] = src_tokens.eq(cfg.padding_idx)
        if patch_masks is not None and not artificial:
            pad[:, :P] = (~patch_masks.to(dev).bool()).unsqueeze(1)
        if has_pads is None:
            has_pads = bool(pad.any())  # same host sync as encoder_module.py:742
        kpm = pad if has_pads else None

        biases, pos = self._encoder_bias(h, w, T_txt, artificial)
        L0 = self.enc_layers[0]
        x = torch.empty((B * T, D), dtype=torch.float32, device=dev)  # fp32 residual stream
        a = torch.empty((B * T, D), dtype=_BF16, device=dev)
        # image rows: image_proj -> +type_embedding(1) -> patch_layernorm_embedding  (:416-423)
        if artificial:
            proj = bag
        else:
            proj = ops.gemm(feat.view(B * P, 1024), self.w_image_proj, bias=self.b_image_proj, out_dtype=torch.float32)
        ops.row_layernorm(proj, pre_add=self.type_img, ln1=self.ln_patch, out1=x, ln2=L0["ln_self"], out2=a,
                          zero_row=pad[:, :P].contiguous().view(-1) if has_pads else None, seg=(P, T, 0))
        # text rows: embed_tokens -> +type_embedding(0) -> layernorm_embedding  (:400-408)
        ops.row_layernorm(self.embed_tokens, rows=B * T_txt, D=D, gather_idx=src_tokens.reshape(-1).contiguous(),
                          pre_add=self.type_txt, ln1=self.ln_emb, out1=x, ln2=L0["ln_self"], out2=a,
                          zero_row=pad[:, P:].contiguous().view(-1) if has_pads else None, seg=(T_txt, T, P))
        a2 = torch.empty_like(a)
        stats = torch.empty((B * T, cfg.ffn_dim // 64, 2), dtype=torch.float32, device=dev)
        for li, L in enumerate(self.enc_layers):
            y = self._self_attention(a, L["attn"], B, T, biases[li], False, kpm)
            ops.row_layernorm(y, ln1=L["ln_attn"], residual=x, out1=x, ln2=L["ln_final"], out2=a2)
            self._ffn(a2, x, L, stats)
            nxt = self.enc_layers[li + 1]["ln_self"] if li + 1 < len(self.enc_layers) else self.ln_enc_out
            ops.row_layernorm(x, ln2=nxt, out2=a)
        return dict(encoder_out=a, B=B, T=T, P=P, hw=(h, w), pad=pad, has_pads=has_pads, pos=pos,
                    image_features=feat, image_proj=proj)

    def encoder_out_dict(self, enc):
        """The dict encoder_module.py:839-851 returns (T x B x C tensors are strided views)."""
        B, T, P, D = enc["B"], enc["T"], enc["P"], self.cfg.embed_dim
        eo = enc["encoder_out"].view(B, T, D)
        return {
            "encoder_out": [eo.transpose(0, 1)],
            "encoder_padding_mask": [enc["pad"].bool()],
            "encoder_embedding": [],
            "encoder_states": [],
            "src_tokens": [],
            "src_lengths": [],
            "position_embeddings": [enc["pos"].unsqueeze(0).expand(B, -1, -1)],
            "patch_images": [],
            "image_embed_before_scale": [enc["image_proj"].view(B, P, D)],
            "image_embed_shape": [enc["hw"]],
            "image_embed_before_proj": [enc["image_features"].view(B, P, -1) if enc["image_features"] is not None else None],
        }

    # ------------------------------------------------------------------------------------
    # decoder
    # ------------------------------------------------------------------------------------
    def decode(self, enc, prev_output_tokens, full_context_alignment=False, features_only=False):
        cfg = self.cfg
        dev = self.device
        D, H = cfg.embed_dim, cfg.heads
        B, Te, P = enc["B"], enc["T"], enc["P"]
        h, w = enc["hw"]
        Td = P + 1
        enc_out = enc["encoder_out"]  # [B*Te, D] bf16
        kpm = enc["pad"] if enc["has_pads"] else None
        self_biases, cross_abs = self._decoder_bias(h, w, enc["pos"])
        L0 = self.dec_layers[0]

        x = torch.empty((B * Td, D), dtype=torch.float32, device=dev)
        a = torch.empty((B * Td, D), dtype=_BF16, device=dev)
        bos = prev_output_tokens.to(dev)[:, 0].contiguous()
        # x = LN_emb(cat([embed_tokens(bos), dec_in]))  (decoder_module.py:530-538, 575-576); embed_scale == 1
        ops.row_layernorm(self.embed_tokens, rows=B, D=D, gather_idx=bos, ln1=self.dec_ln_emb, out1=x,
                          ln2=L0["ln_self"], out2=a, seg=(1, Td, 0))
        idx = self._cached(("dec_in_idx", B, Te, P), lambda: (
            torch.arange(B).unsqueeze(1) * Te + torch.arange(P).unsqueeze(0)).reshape(-1))
        if cfg.decoder_input_type == "encoder_output":
            ops.row_layernorm(enc_out, rows=B * P, gather_idx=idx, ln1=self.dec_ln_emb, out1=x, ln2=L0["ln_self"],
                              out2=a, seg=(P, Td, 1))
        else:  # 'encoder_input': image_embed_before_scale
            ops.row_layernorm(enc["image_proj"], rows=B * P, ln1=self.dec_ln_emb, out1=x, ln2=L0["ln_self"], out2=a,
                              seg=(P, Td, 1))
        # cross-attention K/V of every layer in one GEMM: [B*Te, L*2D]
        nL = len(self.dec_layers)
        kv_all = ops.gemm(enc_out, self.w_cross_kv_all, bias=self.b_cross_kv_all, tag="cross_kv")
        a2 = torch.empty_like(a)
        stats = torch.empty((B * Td, cfg.ffn_dim // 64, 2), dtype=torch.float32, device=dev)
        o = torch.empty((B * Td, D), dtype=_BF16, device=dev)
        for li, L in enumerate(self.dec_layers):
            y = self._self_attention(a, L["attn"], B, Td, self_biases[li], not full_context_alignment, None)
            ops.row_layernorm(y, ln1=L["ln_self_attn"], residual=x, out1=x, ln2=L["ln_enc_attn"], out2=a2)
            C = L["cross"]
            q = ops.gemm(a2, C["wq"], bias=C["bq"], alpha=cfg.attn_scaling, alpha_cols=D, tag="cross_q")
            kbase = kv_all[:, li * 2 * D:]
            ops.attention(q, kbase, kbase[:, D:], o, B=B, H=H, Tq=Td, Tk=Te, q_strides=(D, Td * D),
                          k_strides=(nL * 2 * D, Te * nL * 2 * D), v_strides=(nL * 2 * D, Te * nL * 2 * D),
                          o_strides=(D, Td * D), bias=cross_abs, head_scale=C["c_attn"], key_padding_mask=kpm)
            y = ops.gemm(o, C["wo"], bias=C["bo"], out_dtype=torch.float32, tag="out_proj")
            ops.row_layernorm(y, ln1=L["ln_cross_attn"], residual=x, out1=x, ln2=L["ln_final"], out2=a)
            self._ffn(a, x, L, stats)
            nxt = self.dec_layers[li + 1]["ln_self"] if li + 1 < nL else self.ln_dec_out
            ops.row_layernorm(x, ln2=nxt, out2=a)
        feats = a.view(B, Td, D)
        extra = {"attn": [None], "inner_states": [], "penultimate": feats}
        if features_only:
            return feats, extra
        logits = ops.gemm(a, self.w_seg_proj, out_dtype=torch.float32)  # seg_projection (no bias) :290-294
        return logits.view(B, Td, cfg.num_seg), extra

    # ------------------------------------------------------------------------------------
    # mask: seg_criterion.py:237-244 + :351 (argmax of the x16 bilinear upsample, eos slot dropped)
    # ------------------------------------------------------------------------------------
    def predict_mask(self, logits, hw, out_hw, target=None):
        return ops.upsample_argmax(logits.float().contiguous(), hw[0], hw[1], out_hw[0], out_hw[1], target=target)
