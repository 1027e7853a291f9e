"""TEST INFRASTRUCTURE ONLY -- the travelling oracle for the segofa hot path.

A CPU (plain torch, fp32 by default) restatement of the reference algorithm, written as
flat functions over a *state dict* so that it needs neither fairseq nor /root/reference.
Every function cites the reference lines it follows.  Only tests/, bench.py's
cpu_baseline / --impl reference leg and __graft_entry__.smoke() may import this file; the
product package (ifseg_b200/) never does.

PINNING: oracle/make_golden.py runs this file against the UNMODIFIED reference model
(imported through oracle/ref_shim.py in the build container) on seeded weights/inputs and
asserts agreement to fp32 round-off (<=2e-5 max-abs on logits); the reference's outputs
are committed as tests/golden/*.pt and re-checked by tests/test_oracle_golden.py on any
box.  The reference itself ships no tests/golden vectors for this path, so
those fixtures -- outputs of the reference itself -- are the pin.

`emulate_bf16=True` rounds to bf16 at the points where the reference's `--bf16` run stores
tensors (module outputs), giving the "reference-bf16 noise floor" used to put the CUDA
path's error in context.
"""
import math
from dataclasses import dataclass, field
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F


@dataclass
class SegOFAConfig:
    """Mirror of the arch presets models/segofa/segofa.py:351-467 + shipped flags."""

    embed_dim: int = 768
    ffn_dim: int = 3072
    heads: int = 12
    enc_layers: int = 6
    dec_layers: int = 6
    resnet_blocks: Tuple[int, int, int] = (3, 4, 23)
    num_seg: int = 15
    patch_image_size: int = 480
    orig_patch_image_size: int = 480
    token_bucket_size: int = 256
    image_bucket_size: int = 42
    attn_scale_factor: float = 2.0
    max_source_positions: int = 1024
    decoder_input_type: str = "encoder_output"
    padding_idx: int = 1
    vocab: int = 59458  # len(dict) - num_seg = 59457 + 1 (unify_transformer.py:402)

    @property
    def head_dim(self):
        return self.embed_dim // self.heads

    @staticmethod
    def preset(arch: str, **kw):
        p = {
            "segofa_tiny": dict(embed_dim=256, ffn_dim=1024, heads=4, enc_layers=4, dec_layers=4, resnet_blocks=(3, 4, 6)),
            "segofa_medium": dict(embed_dim=512, ffn_dim=2048, heads=8, enc_layers=4, dec_layers=4, resnet_blocks=(3, 4, 23)),
            "segofa_base": dict(embed_dim=768, ffn_dim=3072, heads=12, enc_layers=6, dec_layers=6, resnet_blocks=(3, 4, 23)),
            "segofa_large": dict(embed_dim=1024, ffn_dim=4096, heads=16, enc_layers=12, dec_layers=12, resnet_blocks=(3, 8, 36)),
            "segofa_huge": dict(embed_dim=1280, ffn_dim=5120, heads=16, enc_layers=24, dec_layers=12, resnet_blocks=(3, 8, 36)),
        }[arch]
        p.update(kw)
        return SegOFAConfig(**p)


# --------------------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------------------
def _r(x, bf16):
    """rounding point of the reference's bf16 run (tensor stored by a module)."""
    return x.to(torch.bfloat16).to(x.dtype) if bf16 else x


def _ln(x, sd, name, bf16=False, eps=1e-5):
    # custom_fairseq/fairseq/modules/layer_norm.py:30-35 -> torch.nn.LayerNorm(eps=1e-5)
    return _r(F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps), bf16)


def _lin(x, sd, name, bf16=False):
    b = sd.get(name + ".bias")
    return _r(F.linear(x, sd[name + ".weight"], b), bf16)


def make_token_bucket_position(bucket_size, max_position=1024):
    # models/segofa/encoder_module.py:71-84
    context_pos = torch.arange(max_position, dtype=torch.long)[:, None]
    memory_pos = torch.arange(max_position, dtype=torch.long)[None, :]
    relative_pos = context_pos - memory_pos
    sign = torch.sign(relative_pos)
    mid = bucket_size // 2
    abs_pos = torch.where((relative_pos < mid) & (relative_pos > -mid), mid - 1, torch.abs(relative_pos))
    log_pos = torch.ceil(torch.log(abs_pos / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    log_pos = log_pos.int()
    bucket_pos = torch.where(abs_pos.le(mid), relative_pos, log_pos * sign).long()
    return bucket_pos + bucket_size - 1


def make_image_bucket_position(bucket_size, num_relative_distance):
    # models/segofa/encoder_module.py:87-104
    coords_h = torch.arange(bucket_size)
    coords_w = torch.arange(bucket_size)
    coords = torch.stack(torch.meshgrid([coords_h, coords_w], indexing="ij"))
    coords_flatten = torch.flatten(coords, 1)
    relative_coords = coords_flatten[:, :, None] - coords_flatten[:, None, :]
    relative_coords = relative_coords.permute(1, 2, 0).contiguous()
    relative_coords[:, :, 0] += bucket_size - 1
    relative_coords[:, :, 1] += bucket_size - 1
    relative_coords[:, :, 0] *= 2 * bucket_size - 1
    idx = torch.zeros(size=(bucket_size * bucket_size + 1,) * 2, dtype=relative_coords.dtype)
    idx[1:, 1:] = relative_coords.sum(-1)
    idx[0, 0:] = num_relative_distance - 3
    idx[0:, 0] = num_relative_distance - 2
    idx[0, 0] = num_relative_distance - 1
    return idx


# --------------------------------------------------------------------------------------
# ResNet stem: models/segofa/resnet.py:215-229, Bottleneck :117-137, frozen_bn.py:36-57
# --------------------------------------------------------------------------------------
def _frozen_bn(x, sd, name, bf16=False, eps=1e-5):
    # frozen_bn.py:40-45 / F.batch_norm(training=False) -- same affine in fp32
    scale = sd[name + ".weight"] * (sd[name + ".running_var"] + eps).rsqrt()
    bias = sd[name + ".bias"] - sd[name + ".running_mean"] * scale
    return _r(x * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1), bf16)


def _bottleneck(x, sd, p, stride, has_down, bf16):
    out = _r(F.conv2d(x, sd[p + ".conv1.weight"]), bf16)
    out = F.relu(_frozen_bn(out, sd, p + ".bn1", bf16))
    out = _r(F.conv2d(out, sd[p + ".conv2.weight"], stride=stride, padding=1), bf16)
    out = F.relu(_frozen_bn(out, sd, p + ".bn2", bf16))
    out = _r(F.conv2d(out, sd[p + ".conv3.weight"]), bf16)
    out = _frozen_bn(out, sd, p + ".bn3", bf16)
    if has_down:
        idn = _r(F.conv2d(x, sd[p + ".downsample.0.weight"], stride=stride), bf16)
        idn = _frozen_bn(idn, sd, p + ".downsample.1", bf16)
    else:
        idn = x
    return F.relu(_r(idn + out, bf16))


def resnet_stem(images, sd, cfg: SegOFAConfig, prefix="encoder.embed_images", bf16=False):
    x = _r(F.conv2d(images, sd[prefix + ".conv1.weight"], stride=2, padding=3), bf16)
    x = F.relu(_frozen_bn(x, sd, prefix + ".bn1", bf16))
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    for li, nblocks in enumerate(cfg.resnet_blocks):
        for bi in range(nblocks):
            stride = 2 if (li > 0 and bi == 0) else 1
            x = _bottleneck(x, sd, f"{prefix}.layer{li + 1}.{bi}", stride, bi == 0, bf16)
    return x  # [B,1024,h,w]


# --------------------------------------------------------------------------------------
# attention: models/segofa/unify_multihead_attention.py:327-523 (bias path)
# --------------------------------------------------------------------------------------
def _mha(xq, xkv, sd, p, cfg, bias, causal=False, key_pad=None, bf16=False, want_probs=False):
    """xq [B,Tq,D], xkv [B,Tk,D], bias [H,Tq,Tk] (batch-invariant) -> [B,Tq,D]."""
    B, Tq, D = xq.shape
    Tk = xkv.shape[1]
    H, dh = cfg.heads, cfg.head_dim
    scaling = float(dh * cfg.attn_scale_factor) ** -0.5  # :58
    q = _r(_lin(xq, sd, p + ".q_proj", bf16) * scaling, bf16)  # :328-346
    k = _lin(xkv, sd, p + ".k_proj", bf16)
    v = _lin(xkv, sd, p + ".v_proj", bf16)
    q = q.view(B, Tq, H, dh).transpose(1, 2)
    k = k.view(B, Tk, H, dh).transpose(1, 2)
    v = v.view(B, Tk, H, dh).transpose(1, 2)
    s = _r(torch.matmul(q, k.transpose(-1, -2)), bf16)  # :459
    s = _r(s + _r(bias, bf16).unsqueeze(0), bf16)  # :464-465
    if causal:  # :467-471 with decoder_module.py:878-890 (triu of -inf)
        mask = torch.full((Tq, Tk), float("-inf"), dtype=s.dtype, device=s.device).triu(1)
        s = s + mask
    if key_pad is not None:  # :477-489
        s = s.masked_fill(key_pad[:, None, None, :], float("-inf"))
    pr = F.softmax(s.float(), dim=-1)  # :494-497 (fp32 softmax)
    o = _r(torch.matmul(_r(pr.to(s.dtype), bf16), v), bf16)  # :501
    o = o * sd[p + ".c_attn"].view(1, H, 1, 1)  # :509-512
    o = _r(o, bf16).transpose(1, 2).reshape(B, Tq, D)
    o = _lin(o, sd, p + ".out_proj", bf16)  # :513
    return (o, pr) if want_probs else o


def _ffn(x, sd, p, bf16):
    # unify_transformer_layer.py:276-284 ; gelu in fp32 (custom_fairseq/fairseq/modules/gelu.py:24-25)
    h = _lin(x, sd, p + ".fc1", bf16)
    h = _r(F.gelu(h.float()).to(h.dtype), bf16)
    h = _ln(h, sd, p + ".ffn_layernorm", bf16)
    return _lin(h, sd, p + ".fc2", bf16)


def encoder_layer(x, sd, p, cfg, bias, key_pad, bf16=False):
    # unify_transformer_layer.py:222-292 (pre-LN; eval => dropout/DropPath identity)
    r = x
    y = _ln(x, sd, p + ".self_attn_layer_norm", bf16)
    y = _mha(y, y, sd, p + ".self_attn", cfg, bias, key_pad=key_pad, bf16=bf16)
    y = _ln(y, sd, p + ".attn_ln", bf16)
    x = _r(r + y, bf16)
    r = x
    y = _ln(x, sd, p + ".final_layer_norm", bf16)
    y = _ffn(y, sd, p, bf16)
    return _r(r + y, bf16)


def decoder_layer(x, enc, sd, p, cfg, self_bias, cross_bias, causal, enc_pad, bf16=False, want_probs=False):
    # unify_transformer_layer.py:431-581
    r = x
    y = _ln(x, sd, p + ".self_attn_layer_norm", bf16)
    y = _mha(y, y, sd, p + ".self_attn", cfg, self_bias, causal=causal, bf16=bf16)
    y = _ln(y, sd, p + ".self_attn_ln", bf16)
    x = _r(r + y, bf16)
    r = x
    y = _ln(x, sd, p + ".encoder_attn_layer_norm", bf16)
    res = _mha(y, enc, sd, p + ".encoder_attn", cfg, cross_bias, key_pad=enc_pad, bf16=bf16, want_probs=want_probs)
    y, probs = res if want_probs else (res, None)
    y = _ln(y, sd, p + ".cross_attn_ln", bf16)
    x = _r(r + y, bf16)
    r = x
    y = _ln(x, sd, p + ".final_layer_norm", bf16)
    y = _ffn(y, sd, p, bf16)
    return _r(r + y, bf16), probs


# --------------------------------------------------------------------------------------
# position bias
# --------------------------------------------------------------------------------------
def _image_position_ids(h, w, bucket):
    # encoder_module.py:339-341
    return (torch.arange(w).unsqueeze(0).expand(h, w) + torch.arange(h).unsqueeze(1) * bucket + 1).reshape(-1)


def _interp_grid(t, src_hw, dst_hw):
    """t [..., C, hs, ws] -> bilinear(align_corners=False) -> [..., C, hd, wd]"""
    return F.interpolate(t, size=dst_hw, mode="bilinear")


def encoder_image_pos_embed(sd, cfg, h, w, device):
    # encoder_module.py:358-370 ; returns [P,D] (batch-invariant)
    P = h * w
    orig_hw = cfg.orig_patch_image_size // 16
    tab = sd["encoder.embed_image_positions.weight"]
    if P > orig_hw * orig_hw:
        ids = _image_position_ids(orig_hw, orig_hw, cfg.image_bucket_size).to(device)
        old = tab[ids].reshape(1, orig_hw, orig_hw, -1).permute(0, 3, 1, 2)
        new = _interp_grid(old, (orig_hw, orig_hw), (h, w))
        return new.permute(0, 2, 3, 1).reshape(P, -1)
    ids = _image_position_ids(h, w, cfg.image_bucket_size).to(device)
    return tab[ids]


def encoder_rel_bias(sd, cfg, layer, h, w, T_txt, artificial=False):
    """rel-pos part of encoder self-attn bias [H,T_e,T_e]; encoder_module.py:313-331, 790-808
    (real image: image block gathered on the ORIG grid and interpolated key axis first, then
    query axis) / :631-635 (artificial image: actual-grid ids, no interpolation)."""
    H = cfg.heads
    P = h * w
    T = P + T_txt
    dev = sd["encoder.token_rp_bucket"].device
    out = torch.zeros(H, T, T, dtype=sd[f"encoder.token_rel_pos_table_list.{layer}.weight"].dtype, device=dev)
    rp = sd["encoder.token_rp_bucket"][:T_txt, :T_txt]
    out[:, P:, P:] = sd[f"encoder.token_rel_pos_table_list.{layer}.weight"][rp].permute(2, 0, 1)
    itab = sd[f"encoder.image_rel_pos_table_list.{layer}.weight"]
    ibucket = sd["encoder.image_rp_bucket"]
    if artificial:
        ids = _image_position_ids(h, w, cfg.image_bucket_size).to(dev)
        out[:, :P, :P] = itab[ibucket[ids][:, ids]].permute(2, 0, 1)
        return out
    oh = cfg.orig_patch_image_size // 16
    ids = _image_position_ids(oh, oh, cfg.image_bucket_size).to(dev)
    v = itab[ibucket[ids][:, ids]].permute(2, 0, 1)  # [H, oh*oh (q), oh*oh (k)]
    # 'b d (h1 w1) (h2 w2) -> (b h1 w1) d h2 w2' ; interpolate over the key grid
    v = v.reshape(H, oh * oh, oh, oh).permute(1, 0, 2, 3)
    v = _interp_grid(v, (oh, oh), (h, w))  # [(h1 w1), H, h, w]
    # '(b h1 w1) d h2 w2 -> (b h2 w2) d h1 w1' ; interpolate over the query grid
    v = v.reshape(oh, oh, H, h * w).permute(3, 2, 0, 1)
    v = _interp_grid(v, (oh, oh), (h, w))  # [(h2 w2), H, h, w]
    out[:, :P, :P] = v.reshape(h * w, H, h * w).permute(1, 2, 0)
    return out


def decoder_seg_pos_embed(sd, cfg, h, w):
    # decoder_module.py:541-550 ; [T_d, D]
    sb = cfg.patch_image_size // 16
    tab = sd["decoder.embed_seg_positions.weight"]
    ids = (torch.arange(sb).unsqueeze(0).expand(sb, sb) + torch.arange(sb).unsqueeze(1) * sb + 1).to(tab.device)
    old = tab[ids].reshape(1, sb, sb, -1).permute(0, 3, 1, 2)
    new = _interp_grid(old, (sb, sb), (h, w)).permute(0, 2, 3, 1).reshape(h * w, -1)
    return torch.cat([tab[0:1], new], dim=0)


def decoder_seg_rel_bias(sd, cfg, layer, h, w):
    """decoder_module.py:601-625: seg rel-pos table on the seg_bucket grid, interpolated
    (query axis first, then key axis) with the bos row/column carried through."""
    H = cfg.heads
    sb = cfg.patch_image_size // 16
    n = sb * sb
    Td = h * w + 1
    v = sd[f"decoder.seg_rel_pos_table_list.{layer}.weight"][sd["decoder.seg_rp_bucket"]].permute(2, 0, 1)  # [H,n+1(q),n+1(k)]
    # 'b c hw1 hw2 -> (b hw2) c hw1': batch = key index, interpolate the query axis
    t = v.permute(2, 0, 1)  # [k, H, q]
    bos, seg = t[..., :1], t[..., 1:]
    seg = _interp_grid(seg.reshape(n + 1, H, sb, sb), (sb, sb), (h, w)).reshape(n + 1, H, h * w)
    t = torch.cat([bos, seg], dim=-1)  # [k(n+1), H, q(Td)]
    # '(b hw2) c hw1 -> (b hw1) c hw2': batch = query index, interpolate the key axis
    t = t.permute(2, 1, 0)  # [q(Td), H, k(n+1)]
    bos, seg = t[..., :1], t[..., 1:]
    seg = _interp_grid(seg.reshape(Td, H, sb, sb), (sb, sb), (h, w)).reshape(Td, H, h * w)
    t = torch.cat([bos, seg], dim=-1)  # [q, H, k(Td)]
    return t.permute(1, 0, 2).contiguous()  # [H, Td, Td]


def _abs_bias(pos_q_in, pos_k_in, sd, qname, kname, cfg, bf16):
    # encoder_module.py:765-771 / decoder_module.py:335-366
    H, dh = cfg.heads, cfg.head_dim
    pos_scaling = float(cfg.embed_dim / cfg.heads * cfg.attn_scale_factor) ** -0.5
    pq = _r(_lin(pos_q_in, sd, qname, bf16).view(-1, H, dh).transpose(0, 1) * pos_scaling, bf16)
    pk = _lin(pos_k_in, sd, kname, bf16).view(-1, H, dh).transpose(0, 1)
    return _r(torch.matmul(pq, pk.transpose(1, 2)), bf16)  # [H,Tq,Tk]


# --------------------------------------------------------------------------------------
# encoder: encoder_module.py:677-851 (real image) and :499-675 (artificial image)
# --------------------------------------------------------------------------------------
def encode(sd, cfg: SegOFAConfig, src_tokens, patch_images=None, patch_masks=None, bf16=False,
           bag_tokens=None, bag_offsets=None, image_features=None):
    """Returns dict(encoder_out [B,T_e,D], position_embeddings [T_e,D], image_embed_shape,
    image_embed_before_proj [B,P,1024], image_embed_before_scale [B,P,D], encoder_padding_mask).
    `image_features` short-circuits the stem (for layer-level tests)."""
    B, T_txt = src_tokens.shape
    dev = src_tokens.device
    artificial = bag_tokens is not None
    if artificial:
        # encoder_module.py:529-551 -- ragged EmbeddingBag(mean) per patch
        h = w = cfg.patch_image_size // 16
        tokens = bag_tokens[bag_tokens != cfg.padding_idx]
        off = bag_offsets.view(B, -1)
        off = torch.cat([off.new_zeros(B, 1), off], dim=1)
        base = torch.cat([off.new_zeros(1), off[:-1, -1]]).cumsum(0)
        off = (off + base.unsqueeze(1))[:, :-1].flatten()
        img_x = F.embedding_bag(tokens, sd["encoder.embed_tokens.weight"], off, mode="mean").view(B, h * w, -1)
        feat_seq = None
        before_scale = img_x
    else:
        feat = image_features if image_features is not None else resnet_stem(patch_images, sd, cfg, bf16=bf16)
        h, w = feat.shape[-2:]
        feat_seq = feat.flatten(2).transpose(1, 2)  # [B,P,1024] :344
        before_scale = _lin(feat_seq, sd, "encoder.image_proj", bf16)  # :416
        img_x = before_scale
    P = h * w
    image_pad = torch.zeros(B, P, dtype=torch.bool, device=dev)
    if patch_masks is not None and not artificial:
        image_pad[~patch_masks] = True  # :730
    pad = torch.cat([image_pad, src_tokens.eq(cfg.padding_idx)], dim=1)  # :737-739
    has_pads = bool(pad.any())

    type_emb = sd["encoder.type_embedding.weight"]
    tok = sd["encoder.embed_tokens.weight"][src_tokens]  # embed_scale == 1 (no_scale_embedding)
    x_txt = _ln(_r(tok + type_emb[0], bf16), sd, "encoder.layernorm_embedding", bf16)  # :402-408
    x_img = _ln(_r(img_x + type_emb[1], bf16), sd, "encoder.patch_layernorm_embedding", bf16)  # :417-423
    x = torch.cat([x_img, x_txt], dim=1)  # :427
    if has_pads:
        x = x * (1 - pad.unsqueeze(-1).type_as(x))  # :751-752

    pos_txt = _ln(sd["encoder.embed_positions.weight"][:T_txt], sd, "encoder.pos_ln", bf16)  # :744,757
    pos_img = _ln(encoder_image_pos_embed(sd, cfg, h, w, dev), sd, "encoder.image_pos_ln", bf16)
    pos = torch.cat([pos_img, pos_txt], dim=0)  # [T_e,D] :760
    abs_bias = _abs_bias(pos, pos, sd, "encoder.pos_q_linear", "encoder.pos_k_linear", cfg, bf16)

    for l in range(cfg.enc_layers):
        bias = _r(abs_bias + encoder_rel_bias(sd, cfg, l, h, w, T_txt, artificial), bf16)  # :790-808
        x = encoder_layer(x, sd, f"encoder.layers.{l}", cfg, bias, pad if has_pads else None, bf16)
    x = _ln(x, sd, "encoder.layer_norm", bf16)  # :829-830
    return dict(encoder_out=x, position_embeddings=pos, image_embed_shape=(h, w), encoder_padding_mask=pad,
                image_embed_before_proj=feat_seq, image_embed_before_scale=before_scale)


# --------------------------------------------------------------------------------------
# decoder: decoder_module.py:486-677 (surrogate) + output_projection :290-294
# --------------------------------------------------------------------------------------
def decode(sd, cfg: SegOFAConfig, enc: Dict, prev_output_tokens, full_context_alignment=False, bf16=False,
           want_attn=False):
    x_enc = enc["encoder_out"]
    B = x_enc.shape[0]
    h, w = enc["image_embed_shape"]
    P = h * w
    bos = sd["decoder.embed_tokens.weight"][prev_output_tokens[:, :1]]  # :530,537
    dec_in = x_enc[:, :P] if cfg.decoder_input_type == "encoder_output" else enc["image_embed_before_scale"]
    x = torch.cat([bos, dec_in], dim=1)  # embed_scale == 1
    tgt_pos = decoder_seg_pos_embed(sd, cfg, h, w)  # [T_d,D]
    tgt_pos_n = _ln(tgt_pos, sd, "decoder.seg_pos_ln", bf16)
    self_abs = _abs_bias(tgt_pos_n, tgt_pos_n, sd, "decoder.self_pos_q_linear", "decoder.self_pos_k_linear", cfg, bf16)
    cross_abs = _abs_bias(tgt_pos_n, enc["position_embeddings"], sd, "decoder.cross_pos_q_linear",
                          "decoder.cross_pos_k_linear", cfg, bf16)
    x = _ln(x, sd, "decoder.layernorm_embedding", bf16)  # :575-576 (disable_entangle => no pos add)
    # the decoder ALWAYS passes the encoder padding mask (:522-523, 647)
    enc_pad = enc["encoder_padding_mask"]
    enc_pad = enc_pad if bool(enc_pad.any()) else None  # all-False mask is a no-op
    probs = None
    for l in range(cfg.dec_layers):
        self_bias = _r(self_abs + decoder_seg_rel_bias(sd, cfg, l, h, w), bf16)  # :601-627
        last = l == cfg.dec_layers - 1
        x, pr = decoder_layer(x, x_enc, sd, f"decoder.layers.{l}", cfg, self_bias, cross_abs,
                              causal=not full_context_alignment, enc_pad=enc_pad, bf16=bf16,
                              want_probs=want_attn and last)
        probs = pr if pr is not None else probs
    x = _ln(x, sd, "decoder.layer_norm", bf16)  # :668-669
    logits = _r(F.linear(x, sd["decoder.seg_projection.weight"]), bf16)  # :290-294
    extra = {"penultimate": x}
    if probs is not None:
        extra["attn"] = probs.mean(dim=1)  # :661-666 average over heads -> [B,T_d,T_e]
    return logits, extra


def segofa_forward(sd, cfg, src_tokens, patch_images, patch_masks=None, prev_output_tokens=None,
                   full_context_alignment=False, bf16=False, want_attn=False):
    """models/segofa/segofa.py:69-134 (real-image branch) -> logits [B,P+1,C], extra."""
    B = src_tokens.shape[0]
    if prev_output_tokens is None:
        prev_output_tokens = torch.zeros(B, 1, dtype=torch.long, device=src_tokens.device)
    enc = encode(sd, cfg, src_tokens, patch_images, patch_masks, bf16=bf16)
    logits, extra = decode(sd, cfg, enc, prev_output_tokens, full_context_alignment, bf16, want_attn)
    extra["encoder_returns"] = enc
    return logits, extra


def segofa_forward_aux(sd, cfg, aux_input, bf16=False):
    """models/segofa/segofa.py:136-151 (image-free branch; always causal)."""
    enc = encode(sd, cfg, aux_input["src_tokens"], bf16=bf16, bag_tokens=aux_input["patch_images"],
                 bag_offsets=aux_input["patch_masks"])
    return decode(sd, cfg, enc, aux_input["prev_output_tokens"], False, bf16)


# --------------------------------------------------------------------------------------
# criterion pieces: criterions/seg_criterion.py:237-267, 349-362
# --------------------------------------------------------------------------------------
def upsample_logits(logits, hp, wp, h, w):
    """seg_criterion.py:237-244 with mmseg.ops.resize == F.interpolate(bilinear, align_corners=False)
    (mmsegmentation v0.28.0 mmseg/ops/wrappers.py; not vendored).  logits [B,P+1,C] -> [B,h*w+1,C]."""
    B, _, C = logits.shape
    x = logits[:, :-1].reshape(B, hp, wp, C).permute(0, 3, 1, 2)
    x = F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)
    x = x.permute(0, 2, 3, 1).reshape(B, h * w, C)
    return torch.cat([x, logits[:, -1:]], dim=1)


def predict_mask(logits, hp, wp, h, w):
    """argmax mask of compute_metric (seg_criterion.py:294,299,351): [B,h*w] int64."""
    return upsample_logits(logits.float(), hp, wp, h, w)[:, :-1].argmax(-1)


def compute_metric(scores, target):
    # seg_criterion.py:349-362
    C = scores.size(-1)
    pred = scores.argmax(-1)
    inter = pred[pred == target]
    ai = torch.histc(inter.float(), bins=C, min=0, max=C - 1)
    ap = torch.histc(pred.float(), bins=C, min=0, max=C - 1)
    al = torch.histc(target.float(), bins=C, min=0, max=C - 1)
    return ai, ap, al, ap + al - ai


def imfree_loss(logits, target, cfg: SegOFAConfig, seg_id_offset=59457, label_smoothing=0.0):
    """seg_criterion.py:246-267 with the hard-coded 32/512 replaced by patch_image_size
    (SURVEY.md s7 'reference quirks'): target [B,S*S+1] of dictionary ids."""
    S = cfg.patch_image_size
    lg = upsample_logits(logits.float(), S // 16, S // 16, S, S)[:, :-1]
    tg = target[:, :-1]
    lg = lg.reshape(-1, lg.size(-1))
    tg = tg.reshape(-1)
    mask = torch.logical_and(tg != cfg.padding_idx, tg != (seg_id_offset + cfg.num_seg))
    return F.cross_entropy(lg[mask], tg[mask] - seg_id_offset, label_smoothing=label_smoothing)
