"""TEST INFRASTRUCTURE ONLY -- the rounding-matched mode of the travelling oracle.

oracle/restated.py restates the reference algorithm in fp32 (the TRUTH, pinned on the unmodified
reference through tests/golden/).  north_star's tolerance -- "<=1e-3 rel in bf16 logits" -- is only
meaningful between two runs that quantise at the same places: the reference's own bf16 run differs
from its fp32 run by 1.4-1.6e-2.  This file is the same algorithm (it calls restated.py for every
index table / relative-position block / interpolation) evaluated on CPU in fp32 with bf16 / fp16
quantisation inserted EXACTLY where the CUDA engine (ifseg_b200/engine.py + csrc/) stores a tensor
in reduced precision:

  stem            NHWC bf16 activations; bf16 weights; epilogue (acc*scale + bias (+idn)) -> relu -> bf16
  LayerNorms      fp32 statistics on the fp32 residual stream; the GEMM-operand copy is bf16
  linears         bf16 x bf16 -> fp32 accumulate; q/k/v, cross-q, cross-k/v, fc1(+GELU) outputs bf16;
                  out_proj / fc2 / image_proj / seg_projection outputs fp32
  ffn_layernorm   folded into fc2 (inference engine): statistics of the bf16 GELU output,
                  W2' = bf16(W2 * gamma), x += rstd * (f W2'^T - mean * rowsum(W2')) + (b2 + W2 beta)
                  (`fold_ffn=False` = the training engine: bf16 pre-activation h, z = bf16(LN(gelu(h))), fc2 fp32)
  position bias   bf16 position embeddings / projections, fp32 per-head dot product, + rel table, FP16 storage
  attention       S = q k^T fp32 + fp16 bias; the probabilities of every 64-key tile are expressed against the
                  row's reference max m_used (max of the first tile, replaced only when a later tile exceeds it by
                  more than 2^8 -- csrc/attention.cu kRescaleThreshold), P = bf16(exp2(...)), row sum from the
                  unrounded fp32 p, O = P V fp32, o = bf16(O * (1/l * c_attn))

`quantize=False` turns every quantiser into the identity; tests/test_oracle_golden.py checks that this
mode reproduces restated.py (hence the reference) to fp32 round-off, i.e. the file is the reference
algorithm and nothing else.  Only tests/, bench.py's parity leg and __graft_entry__.smoke() import it.
"""
import torch
import torch.nn.functional as F

from . import restated as R

_LOG2E = 1.4426950408889634
_RESCALE_LOG2 = 8.0  # csrc/attention.cu: kRescaleThreshold
_KTILE = 64          # csrc/attention.cu: kKTile


class Q:
    """Quantisers of one run.  `jitter` > 0 multiplies every value by (1 + jitter * N(0,1)) right before it is rounded --
    a model of a different fp32 accumulation order (jitter = 1e-7 is about one fp32 ulp).  A quantised deep chain is
    chaotic in this perturbation: a value that crosses a rounding boundary moves by a whole bf16 ulp, so a relative
    perturbation d becomes sqrt(d * ulp) after one storage point and saturates at the quantisation-noise level after a
    handful of them.  The distance between a jittered and an unjittered run is therefore the best agreement ANY two
    implementations that are not bit-identical in their accumulation order can have (tests/test_oracle_golden.py)."""

    def __init__(self, quantize=True, jitter=0.0, seed=0):
        self.on = quantize
        self.jitter = jitter
        self.gen = torch.Generator().manual_seed(seed) if jitter > 0 else None

    def _j(self, x):
        if self.jitter > 0:
            return x * (1.0 + self.jitter * torch.randn(x.shape, generator=self.gen))
        return x

    def bf(self, x):
        return self._j(x).to(torch.bfloat16).to(torch.float32) if self.on else x

    def h(self, x):
        return self._j(x).to(torch.float16).to(torch.float32) if self.on else x


def _ln(x, sd, name, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _lin(x, sd, name, q: Q, bias=True):
    """bf16 operands, fp32 accumulate, fp32 bias; the caller decides how the output is stored."""
    y = q.bf(x) @ q.bf(sd[name + ".weight"]).t()
    return y + sd[name + ".bias"] if bias else y


# --------------------------------------------------------------------------------------
# stem (resnet.py:215-229, frozen_bn.py:40-45) with the engine's storage points
# --------------------------------------------------------------------------------------
def _bn_affine(sd, name, eps=1e-5):
    scale = sd[name + ".weight"] * (sd[name + ".running_var"] + eps).rsqrt()
    return scale.view(1, -1, 1, 1), (sd[name + ".bias"] - sd[name + ".running_mean"] * scale).view(1, -1, 1, 1)


def _conv_bn(x, sd, conv, bn, q: Q, stride=1, padding=0, relu=True, residual=None):
    s, b = _bn_affine(sd, bn)
    y = F.conv2d(x, q.bf(sd[conv + ".weight"]), stride=stride, padding=padding) * s + b
    if residual is not None:
        y = y + residual
    return q.bf(F.relu(y) if relu else y)


def resnet_stem(images, sd, cfg, q: Q, prefix="encoder.embed_images"):
    x = _conv_bn(q.bf(images), sd, prefix + ".conv1", prefix + ".bn1", q, stride=2, padding=3)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nblocks in enumerate(cfg.resnet_blocks):
        for bi in range(nblocks):
            p = f"{prefix}.layer{li + 1}.{bi}"
            stride = 2 if (li > 0 and bi == 0) else 1
            idn = x if bi != 0 else _conv_bn(x, sd, p + ".downsample.0", p + ".downsample.1", q, stride=stride, relu=False)
            y = _conv_bn(x, sd, p + ".conv1", p + ".bn1", q)
            y = _conv_bn(y, sd, p + ".conv2", p + ".bn2", q, stride=stride, padding=1)
            x = _conv_bn(y, sd, p + ".conv3", p + ".bn3", q, residual=idn)  # relu(bn3(conv3) + identity) :128-135
    return x


# --------------------------------------------------------------------------------------
# attention (unify_multihead_attention.py:459-512) with the kernel's tile-wise reference max
# --------------------------------------------------------------------------------------
def attention(qh, kh, vh, bias, c_attn, q: Q, causal=False, key_pad=None):
    """qh [B,H,Tq,dh] (pre-scaled), kh/vh [B,H,Tk,dh]: bf16-representable fp32; bias [H,Tq,Tk] (fp16-representable
    when quantising) -> [B,Tq,H*dh] fp32, bf16-representable."""
    B, H, Tq, dh = qh.shape
    Tk = kh.shape[2]
    s = qh @ kh.transpose(-1, -2) + bias.unsqueeze(0)
    if causal:
        s = s + torch.full((Tq, Tk), float("-inf")).triu(1)
    if key_pad is not None:
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    if not q.on:
        p = F.softmax(s, dim=-1)
        o = (p @ vh) * c_attn.view(1, H, 1, 1)
    else:
        nt = (Tk + _KTILE - 1) // _KTILE
        sp = F.pad(s, (0, nt * _KTILE - Tk), value=float("-inf")).view(B, H, Tq, nt, _KTILE)
        tmax = sp.amax(-1) * _LOG2E                                      # [B,H,Tq,nt] log2-domain tile maxima
        m = torch.where(torch.isinf(tmax[..., 0]), torch.zeros_like(tmax[..., 0]), tmax[..., 0])
        m_used = [m]
        for j in range(1, nt):                                           # lazy rescale rule of the kernel
            m = torch.where(tmax[..., j] > m + _RESCALE_LOG2, tmax[..., j], m)
            m_used.append(m)
        m_used = torch.stack(m_used, dim=-1)                             # [B,H,Tq,nt]
        p = torch.exp2(sp * _LOG2E - m_used.unsqueeze(-1))               # fp32, relative to the tile's reference max
        scale = torch.exp2(m_used - m_used[..., -1:]).unsqueeze(-1)      # what the lazy rescales multiply in later
        l = (p * scale).sum(dim=(-1, -2))
        pe = (q.bf(p) * scale).view(B, H, Tq, nt * _KTILE)[..., :Tk]
        o = (pe @ vh) * ((1.0 / l) * c_attn.view(1, H, 1)).unsqueeze(-1)  # inv = (1/l) * head_scale, one multiply
    return q.bf(o).transpose(1, 2).reshape(B, Tq, H * dh)


def _split(x, H):
    B, T, D = x.shape
    return x.view(B, T, H, D // H).transpose(1, 2)


def _self_attn(a, sd, p, cfg, bias, q: Q, causal=False, key_pad=None):
    sc = float(cfg.head_dim * cfg.attn_scale_factor) ** -0.5
    H = cfg.heads
    qq = q.bf(_lin(a, sd, p + ".q_proj", q) * sc)
    kk = q.bf(_lin(a, sd, p + ".k_proj", q))
    vv = q.bf(_lin(a, sd, p + ".v_proj", q))
    o = attention(_split(qq, H), _split(kk, H), _split(vv, H), bias, sd[p + ".c_attn"], q, causal, key_pad)
    return _lin(o, sd, p + ".out_proj", q)  # fp32


def _cross_attn(a, kk, vv, sd, p, cfg, bias, q: Q, key_pad=None):
    sc = float(cfg.head_dim * cfg.attn_scale_factor) ** -0.5
    H = cfg.heads
    qq = q.bf(_lin(a, sd, p + ".q_proj", q) * sc)
    o = attention(_split(qq, H), _split(kk, H), _split(vv, H), bias, sd[p + ".c_attn"], q, False, key_pad)
    return _lin(o, sd, p + ".out_proj", q)


def _ffn(a, x, sd, p, cfg, q: Q, fold_ffn=True):
    """x + fc2(ffn_layernorm(gelu(fc1(a)))) (unify_transformer_layer.py:276-291) -> fp32 residual stream."""
    ln = p + ".ffn_layernorm"
    if fold_ffn and q.on:
        f = q.bf(F.gelu(_lin(a, sd, p + ".fc1", q)))
        Fd = f.shape[-1]
        mean = f.sum(-1, keepdim=True) / Fd
        var = ((f * f).sum(-1, keepdim=True) / Fd - mean * mean).clamp_min(0.0)
        rstd = torch.rsqrt(var + 1e-5)
        w2 = sd[p + ".fc2.weight"]
        w2f = q.bf(w2 * sd[ln + ".weight"].unsqueeze(0))
        u = w2f.sum(dim=1)
        c = sd[p + ".fc2.bias"] + w2 @ sd[ln + ".bias"]
        return x + (rstd * (f @ w2f.t() - mean * u) + c)
    h = q.bf(_lin(a, sd, p + ".fc1", q))                 # training engine: the saved pre-activation is bf16
    z = q.bf(_ln(F.gelu(h), sd, ln))
    return x + _lin(z, sd, p + ".fc2", q)


# --------------------------------------------------------------------------------------
# position bias (encoder_module.py:757-809, decoder_module.py:541-629)
# --------------------------------------------------------------------------------------
def _abs_bias(pos_q, pos_k, sd, qname, kname, cfg, q: Q):
    H, dh = cfg.heads, cfg.head_dim
    sc = float(cfg.embed_dim / cfg.heads * cfg.attn_scale_factor) ** -0.5
    pq = q.bf(_lin(pos_q, sd, qname, q) * sc).view(-1, H, dh).transpose(0, 1)
    pk = q.bf(_lin(pos_k, sd, kname, q)).view(-1, H, dh).transpose(0, 1)
    return pq @ pk.transpose(1, 2)  # fp32 [H,Tq,Tk]


# --------------------------------------------------------------------------------------
# encoder / decoder
# --------------------------------------------------------------------------------------
def encode(sd, cfg, src_tokens, patch_images=None, patch_masks=None, q: Q = None, bag_tokens=None, bag_offsets=None,
           fold_ffn=True):
    q = q or Q()
    B, T_txt = src_tokens.shape
    dev = src_tokens.device
    artificial = bag_tokens is not None
    feat_seq = None
    if artificial:  # encoder_module.py:529-551
        h = w = cfg.patch_image_size // 16
        tokens = bag_tokens[bag_tokens != cfg.padding_idx]
        off = bag_offsets.view(B, -1)
        off = torch.cat([off.new_zeros(B, 1), off], dim=1)
        base = torch.cat([off.new_zeros(1), off[:-1, -1]]).cumsum(0)
        off = (off + base.unsqueeze(1))[:, :-1].flatten()
        proj = F.embedding_bag(tokens, sd["encoder.embed_tokens.weight"], off, mode="mean").view(B, h * w, -1)
    else:
        feat = resnet_stem(patch_images, sd, cfg, q)
        h, w = feat.shape[-2:]
        feat_seq = feat.flatten(2).transpose(1, 2)
        proj = _lin(feat_seq, sd, "encoder.image_proj", q)  # fp32
    P = h * w
    image_pad = torch.zeros(B, P, dtype=torch.bool, device=dev)
    if patch_masks is not None and not artificial:
        image_pad[~patch_masks] = True
    pad = torch.cat([image_pad, src_tokens.eq(cfg.padding_idx)], dim=1)
    has_pads = bool(pad.any())
    te = sd["encoder.type_embedding.weight"]
    x_img = _ln(proj + te[1], sd, "encoder.patch_layernorm_embedding")
    x_txt = _ln(sd["encoder.embed_tokens.weight"][src_tokens] + te[0], sd, "encoder.layernorm_embedding")
    x = torch.cat([x_img, x_txt], dim=1)  # fp32 residual stream
    if has_pads:
        x = x * (1 - pad.unsqueeze(-1).type_as(x))

    pos_img = q.bf(_ln(R.encoder_image_pos_embed(sd, cfg, h, w, dev), sd, "encoder.image_pos_ln"))
    pos_txt = q.bf(_ln(sd["encoder.embed_positions.weight"][:T_txt], sd, "encoder.pos_ln"))
    pos = torch.cat([pos_img, pos_txt], dim=0)
    absb = _abs_bias(pos, pos, sd, "encoder.pos_q_linear", "encoder.pos_k_linear", cfg, q)
    kp = pad if has_pads else None
    for l in range(cfg.enc_layers):
        p = f"encoder.layers.{l}"
        bias = q.h(absb + R.encoder_rel_bias(sd, cfg, l, h, w, T_txt, artificial))
        a = q.bf(_ln(x, sd, p + ".self_attn_layer_norm"))
        x = x + _ln(_self_attn(a, sd, p + ".self_attn", cfg, bias, q, key_pad=kp), sd, p + ".attn_ln")
        a = q.bf(_ln(x, sd, p + ".final_layer_norm"))
        x = _ffn(a, x, sd, p, cfg, q, fold_ffn)
    out = q.bf(_ln(x, sd, "encoder.layer_norm"))
    return dict(encoder_out=out, position_embeddings=pos, image_embed_shape=(h, w), encoder_padding_mask=pad,
                image_embed_before_proj=feat_seq, image_embed_before_scale=proj)


def decode(sd, cfg, enc, prev_output_tokens, full_context_alignment=False, q: Q = None, fold_ffn=True):
    q = q or Q()
    x_enc = enc["encoder_out"]
    h, w = enc["image_embed_shape"]
    P = h * w
    H = cfg.heads
    bos = sd["decoder.embed_tokens.weight"][prev_output_tokens[:, :1]]
    dec_in = x_enc[:, :P] if cfg.decoder_input_type == "encoder_output" else enc["image_embed_before_scale"]
    x = _ln(torch.cat([bos, dec_in], dim=1), sd, "decoder.layernorm_embedding")
    tgt_pos = q.bf(_ln(R.decoder_seg_pos_embed(sd, cfg, h, w), sd, "decoder.seg_pos_ln"))
    self_abs = _abs_bias(tgt_pos, tgt_pos, sd, "decoder.self_pos_q_linear", "decoder.self_pos_k_linear", cfg, q)
    cross = q.h(_abs_bias(tgt_pos, enc["position_embeddings"], sd, "decoder.cross_pos_q_linear",
                          "decoder.cross_pos_k_linear", cfg, q))
    enc_pad = enc["encoder_padding_mask"]
    enc_pad = enc_pad if bool(enc_pad.any()) else None
    for l in range(cfg.dec_layers):
        p = f"decoder.layers.{l}"
        self_bias = q.h(self_abs + R.decoder_seg_rel_bias(sd, cfg, l, h, w))
        a = q.bf(_ln(x, sd, p + ".self_attn_layer_norm"))
        x = x + _ln(_self_attn(a, sd, p + ".self_attn", cfg, self_bias, q, causal=not full_context_alignment), sd,
                    p + ".self_attn_ln")
        a = q.bf(_ln(x, sd, p + ".encoder_attn_layer_norm"))
        kk = q.bf(_lin(x_enc, sd, p + ".encoder_attn.k_proj", q))
        vv = q.bf(_lin(x_enc, sd, p + ".encoder_attn.v_proj", q))
        x = x + _ln(_cross_attn(a, kk, vv, sd, p + ".encoder_attn", cfg, cross, q, key_pad=enc_pad), sd,
                    p + ".cross_attn_ln")
        a = q.bf(_ln(x, sd, p + ".final_layer_norm"))
        x = _ffn(a, x, sd, p, cfg, q, fold_ffn)
    feats = q.bf(_ln(x, sd, "decoder.layer_norm"))
    logits = _lin(feats, sd, "decoder.seg_projection", q, bias=False)  # fp32
    return logits, {"penultimate": feats}


def segofa_forward(sd, cfg, src_tokens, patch_images, patch_masks=None, prev_output_tokens=None,
                   full_context_alignment=False, quantize=True, jitter=0.0):
    """models/segofa/segofa.py:69-134 with the CUDA engine's storage points -> (logits [B,P+1,C] fp32, extra)."""
    q = Q(quantize, jitter)
    B = src_tokens.shape[0]
    if prev_output_tokens is None:
        prev_output_tokens = torch.zeros(B, 1, dtype=torch.long, device=src_tokens.device)
    enc = encode(sd, cfg, src_tokens, patch_images, patch_masks, q)
    logits, extra = decode(sd, cfg, enc, prev_output_tokens, full_context_alignment, q)
    extra["encoder_returns"] = enc
    return logits, extra


def segofa_forward_aux(sd, cfg, aux_input, quantize=True, fold_ffn=True, jitter=0.0):
    """models/segofa/segofa.py:136-151 (image-free branch; always causal).  fold_ffn=False = the training engine's
    forward (ifseg_b200/train_engine.py: separate ffn_layernorm row kernel)."""
    q = Q(quantize, jitter)
    enc = encode(sd, cfg, aux_input["src_tokens"], q=q, bag_tokens=aux_input["patch_images"],
                 bag_offsets=aux_input["patch_masks"], fold_ffn=fold_ffn)
    return decode(sd, cfg, enc, aux_input["prev_output_tokens"], False, q, fold_ffn)
