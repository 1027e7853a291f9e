"""Deterministic synthetic weights / inputs for the segofa hot path.

No pretrained checkpoint exists offline, so parity and throughput are
defined on seeded random weights.  The generator is keyed by *parameter name* (not by module
construction order), so the reference model in the build container and this package on the
GPU box obtain bit-identical tensors from `generate_state_dict(cfg, seed)`.

Distributions follow the reference init (BERT-style N(0, 0.02) for Linear/Embedding --
custom_fairseq/fairseq/modules/transformer_sentence_encoder.py:21-53 via segofa.py:32-33,
kaiming fan-out for convolutions -- resnet.py:172-177) but, unlike a fresh init, every
LayerNorm/BN/bias/c_attn parameter is perturbed away from its identity value so that a parity
test exercises all of them, and the frozen-BN statistics keep stem activations O(1).
"""
import math
import zlib
from typing import Dict, List, Tuple

import torch

from .config import SegOFAConfig, image_bucket_position, token_bucket_position


def state_dict_spec(cfg: SegOFAConfig) -> List[Tuple[str, Tuple[int, ...], str]]:
    """(name, shape, kind) for every tensor of the reference state dict, in reference order
    (names: SURVEY.md s8b; 889 tensors for segofa_base)."""
    D, F, H, C = cfg.embed_dim, cfg.ffn_dim, cfg.heads, cfg.num_seg
    V = cfg.vocab
    ib2 = cfg.image_bucket_size ** 2 + 1
    tok_rel = 2 * cfg.token_bucket_size - 1
    img_rel = (2 * cfg.image_bucket_size - 1) ** 2 + 3
    sb = cfg.patch_image_size // 16
    seg_rel = (2 * sb - 1) ** 2 + 3
    out = []

    def add(name, shape, kind):
        out.append((name, tuple(shape), kind))

    def ln(name, n):
        add(name + ".weight", (n,), "ln_w")
        add(name + ".bias", (n,), "ln_b")

    def lin(name, o, i, bias=True):
        add(name + ".weight", (o, i), "linear_w")
        if bias:
            add(name + ".bias", (o,), "linear_b")

    def bn(name, n, last=False):
        add(name + ".weight", (n,), "bn_w_last" if last else "bn_w")
        add(name + ".bias", (n,), "bn_b")
        add(name + ".running_mean", (n,), "bn_mean")
        add(name + ".running_var", (n,), "bn_var")

    def mha(name):
        add(name + ".c_attn", (H,), "c_attn")
        for p in ("k_proj", "v_proj", "q_proj", "out_proj"):
            lin(f"{name}.{p}", D, D)

    # ---------------- encoder ----------------
    add("encoder.version", (1,), "version")
    add("encoder.token_rp_bucket", (1024, 1024), "token_rp_bucket")
    add("encoder.image_rp_bucket", (ib2, ib2), "image_rp_bucket")
    add("encoder.embed_tokens.weight", (V, D), "embed_tokens")
    add("encoder.seg_embed_tokens.weight", (C, D), "seg_embed")
    add("encoder.embed_tokens_bag.weight", (V, D), "embed_tokens")
    ln("encoder.layernorm_embedding", D)
    add("encoder.type_embedding.weight", (2, D), "embed")
    r = "encoder.embed_images"
    add(r + ".conv1.weight", (64, 3, 7, 7), "conv")
    bn(r + ".bn1", 64)
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip((64, 128, 256), cfg.resnet_blocks)):
        for bi in range(nblocks):
            p = f"{r}.layer{li + 1}.{bi}"
            add(p + ".conv1.weight", (planes, inplanes, 1, 1), "conv")
            bn(p + ".bn1", planes)
            add(p + ".conv2.weight", (planes, planes, 3, 3), "conv")
            bn(p + ".bn2", planes)
            add(p + ".conv3.weight", (planes * 4, planes, 1, 1), "conv")
            bn(p + ".bn3", planes * 4, last=True)
            if bi == 0:
                add(p + ".downsample.0.weight", (planes * 4, inplanes, 1, 1), "conv")
                bn(p + ".downsample.1", planes * 4)
            inplanes = planes * 4
    lin("encoder.image_proj", D, 1024)
    ln("encoder.patch_layernorm_embedding", D)
    add("encoder.embed_positions.weight", (cfg.max_source_positions + 2, D), "embed")
    add("encoder.embed_image_positions.weight", (ib2, D), "embed")
    ln("encoder.pos_ln", D)
    ln("encoder.image_pos_ln", D)
    lin("encoder.pos_q_linear", D, D)
    lin("encoder.pos_k_linear", D, D)
    for l in range(cfg.enc_layers):
        p = f"encoder.layers.{l}"
        mha(p + ".self_attn")
        ln(p + ".self_attn_layer_norm", D)
        lin(p + ".fc1", F, D)
        lin(p + ".fc2", D, F)
        ln(p + ".attn_ln", D)
        ln(p + ".ffn_layernorm", F)
        ln(p + ".final_layer_norm", D)
    ln("encoder.layer_norm", D)
    for l in range(cfg.enc_layers):
        add(f"encoder.token_rel_pos_table_list.{l}.weight", (tok_rel, H), "rel_table")
    for l in range(cfg.enc_layers):
        add(f"encoder.image_rel_pos_table_list.{l}.weight", (img_rel, H), "rel_table")
    # ---------------- decoder ----------------
    add("decoder.version", (1,), "version")
    add("decoder.seg_rp_bucket", (sb * sb + 1, sb * sb + 1), "seg_rp_bucket")
    add("decoder.token_rp_bucket", (1024, 1024), "token_rp_bucket")
    add("decoder.image_rp_bucket", (ib2, ib2), "image_rp_bucket")
    add("decoder.image_position_idx", (1026,), "image_position_idx")
    add("decoder.bin_id_offset", (1,), "bin_id_offset")
    add("decoder.seg_id_offset", (1,), "seg_id_offset")
    add("decoder.region_prefix", (2,), "region_prefix")
    add("decoder.seg_embed_tokens.weight", (C, D), "seg_embed")
    add("decoder.seg_projection.weight", (C, D), "seg_embed")
    add("decoder.embed_tokens.weight", (V, D), "embed_tokens")
    ln("decoder.layernorm_embedding", D)
    add("decoder.embed_positions.weight", (cfg.max_source_positions + 2, D), "embed")
    add("decoder.embed_image_positions.weight", (ib2, D), "embed")
    add("decoder.embed_seg_positions.weight", (sb * sb + 1, D), "embed")
    ln("decoder.pos_ln", D)
    ln("decoder.image_pos_ln", D)
    ln("decoder.seg_pos_ln", D)
    for n in ("self_pos_q_linear", "self_pos_k_linear", "cross_pos_q_linear", "cross_pos_k_linear"):
        lin("decoder." + n, D, D)
    ln("decoder.code_layernorm_embedding", D)
    for l in range(cfg.dec_layers):
        p = f"decoder.layers.{l}"
        mha(p + ".self_attn")
        ln(p + ".self_attn_ln", D)
        ln(p + ".cross_attn_ln", D)
        ln(p + ".self_attn_layer_norm", D)
        mha(p + ".encoder_attn")
        ln(p + ".encoder_attn_layer_norm", D)
        ln(p + ".ffn_layernorm", F)
        lin(p + ".fc1", F, D)
        lin(p + ".fc2", D, F)
        ln(p + ".final_layer_norm", D)
    ln("decoder.layer_norm", D)
    for l in range(cfg.dec_layers):
        add(f"decoder.token_rel_pos_table_list.{l}.weight", (tok_rel, H), "rel_table")
    for l in range(cfg.dec_layers):
        add(f"decoder.image_rel_pos_table_list.{l}.weight", (img_rel, H), "rel_table")
    for l in range(cfg.dec_layers):
        add(f"decoder.seg_rel_pos_table_list.{l}.weight", (seg_rel, H), "rel_table")
    return out


# names that share one tensor (tied weights under the shipped flags: share_all_embeddings,
# tie_seg_projection=true; unify_transformer.py:336-353, decoder_module.py:134-137)
def _tie_group(name: str) -> str:
    if name in ("encoder.embed_tokens.weight", "encoder.embed_tokens_bag.weight", "decoder.embed_tokens.weight"):
        return "encoder.embed_tokens.weight"
    if name in ("encoder.seg_embed_tokens.weight", "decoder.seg_embed_tokens.weight", "decoder.seg_projection.weight"):
        return "encoder.seg_embed_tokens.weight"
    return name


def _gen(name: str, seed: int) -> torch.Generator:
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)


def generate_state_dict(cfg: SegOFAConfig, seed: int = 0) -> Dict[str, torch.Tensor]:
    """CPU fp32 state dict with the reference's names/shapes (tied entries alias one tensor)."""
    sd: Dict[str, torch.Tensor] = {}
    cache: Dict[str, torch.Tensor] = {}
    sb = cfg.patch_image_size // 16
    for name, shape, kind in state_dict_spec(cfg):
        key = _tie_group(name)
        if key in cache:
            sd[name] = cache[key]
            continue
        g = _gen(key, seed)
        if kind in ("embed", "embed_tokens", "seg_embed", "linear_w", "rel_table"):
            t = torch.randn(shape, generator=g) * 0.02
            if kind == "rel_table":
                t = t * 10.0  # make the rel-pos bias matter (trained tables are O(0.1-1))
            if kind == "embed_tokens":
                t[cfg.padding_idx] = 0
        elif kind == "linear_b":
            t = torch.randn(shape, generator=g) * 0.02
        elif kind == "ln_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            t = 0.05 * torch.randn(shape, generator=g)
        elif kind == "c_attn":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "conv":
            fan_out = shape[0] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * math.sqrt(2.0 / fan_out)
        elif kind == "bn_w":
            t = 0.9 + 0.4 * torch.rand(shape, generator=g)
        elif kind == "bn_w_last":
            t = 0.2 + 0.3 * torch.rand(shape, generator=g)
        elif kind == "bn_b":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_mean":
            t = 0.1 * torch.randn(shape, generator=g)
        elif kind == "bn_var":
            t = 0.6 + 0.8 * torch.rand(shape, generator=g)
        elif kind == "version":
            t = torch.tensor([3.0])
        elif kind == "token_rp_bucket":
            t = token_bucket_position(cfg.token_bucket_size)
        elif kind == "image_rp_bucket":
            t = image_bucket_position(cfg.image_bucket_size, (2 * cfg.image_bucket_size - 1) ** 2 + 3)
        elif kind == "seg_rp_bucket":
            t = image_bucket_position(sb, (2 * sb - 1) ** 2 + 3)
        elif kind == "image_position_idx":
            ws = cfg.code_image_size // 8
            idx = torch.arange(ws).unsqueeze(0).expand(ws, ws) + torch.arange(ws).unsqueeze(1) * cfg.image_bucket_size + 1
            t = torch.cat([torch.tensor([0]), idx.reshape(-1), torch.tensor([1024] * 769)])
        elif kind == "bin_id_offset":
            t = torch.tensor([58457])
        elif kind == "seg_id_offset":
            t = torch.tensor([59457])
        elif kind == "region_prefix":
            t = torch.tensor([976, 35])
        else:
            raise KeyError(kind)
        assert tuple(t.shape) == shape, (name, t.shape, shape)
        cache[key] = t
        sd[name] = t
    return sd


# BPE ids of "what is the segmentation map of the image? object:" + class names (+ "unknown"),
# produced with the reference's utils/BPE files by oracle/make_golden.py and committed in
# tests/golden/prompts.json; bench/tests fall back to random ids of the right LENGTH when the
# fixture is absent (T_txt is what shapes the computation).
PROMPT_LENGTHS = {15: 36, 150: 215, 171: 239}


def synthetic_inputs(cfg: SegOFAConfig, batch: int, image_size: int, seed: int = 1, src_tokens=None):
    """Seeded host inputs of the real-image branch: dict of CPU tensors."""
    g = torch.Generator().manual_seed(seed)
    images = torch.randn(batch, 3, image_size, image_size, generator=g)
    if src_tokens is None:
        t_txt = PROMPT_LENGTHS.get(cfg.num_seg, 20 + cfg.num_seg)
        tok = torch.randint(4, 50000, (t_txt,), generator=g)
        tok[0], tok[-1] = 0, 2
    else:
        tok = torch.as_tensor(src_tokens, dtype=torch.long)
    src = tok.unsqueeze(0).repeat(batch, 1)
    return dict(
        src_tokens=src,
        src_lengths=torch.full((batch,), src.shape[1], dtype=torch.long),
        prev_output_tokens=torch.zeros(batch, 1, dtype=torch.long),
        patch_images=images,
        patch_masks=torch.ones(batch, dtype=torch.bool),
    )


def synthetic_train_sample(cfg: SegOFAConfig, batch: int, image_size: int, seed: int = 1, src_tokens=None,
                           seg_id_offset: int = 59457):
    """Seeded host sample of one image-free training step in the layout SegmentationDataset.collate produces
    (data/mm_data/segmentation_dataset.py:303-347, artificial_image_type='rand_k-1-33'): per sample a random
    sh x sw label grid (sh, sw ~ U{1..32}, labels ~ U{0..C-1}) nearest-resized to the (S/16)^2 patch grid (ragged
    bags of the class-name BPE tokens -> aux_input.patch_images / patch_masks = cumulative bag ends) and to the
    S^2 pixel grid (text2seg_target, dictionary ids + eos); plus the real-image net_input and a random target
    for the no-grad metric pass.  Class-name tokens: stand-in slices of the prompt (1-3 tokens per class)."""
    inp = synthetic_inputs(cfg, batch, image_size, seed, src_tokens)
    g = torch.Generator().manual_seed(seed + 1000)
    C, S = cfg.num_seg, image_size
    hp = S // 16
    tok = inp["src_tokens"][0]
    names = [tok[(13 + n) % (len(tok) - 3): (13 + n) % (len(tok) - 3) + 1 + n % 3] for n in range(C)]
    bags, ends, t2s = [], [], []
    for _ in range(batch):
        sh, sw = (int(torch.randint(1, 33, (1,), generator=g)) for _ in range(2))
        grid = torch.randint(0, C, (1, 1, sh, sw), generator=g).float()
        low = torch.nn.functional.interpolate(grid, size=(hp, hp), mode="nearest").long().reshape(-1)
        full = torch.nn.functional.interpolate(grid, size=(S, S), mode="nearest").long().reshape(-1)
        toks = [names[int(c)] for c in low]
        bags.append(torch.cat(toks))
        ends.append(torch.tensor([len(t) for t in toks]).cumsum(0))
        t2s.append(torch.cat([full + seg_id_offset, torch.tensor([2])]))
    L = max(len(b) for b in bags)
    bag_tokens = torch.full((batch, L), cfg.padding_idx, dtype=torch.long)
    for b in range(batch):
        bag_tokens[b, : len(bags[b])] = bags[b]
    aux = dict(src_tokens=inp["src_tokens"], src_lengths=inp["src_lengths"], patch_images=bag_tokens,
               patch_masks=torch.cat(ends), prev_output_tokens=inp["prev_output_tokens"])
    target = torch.cat([torch.randint(0, C + 1, (batch, S * S), generator=g) + seg_id_offset,
                        torch.full((batch, 1), 2)], 1)
    return dict(net_input=inp, aux_input=aux, target=target, text2seg_target=torch.stack(t2s),
                ntokens=batch * (hp * hp + 1), nsentences=batch)
