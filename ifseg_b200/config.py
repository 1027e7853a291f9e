"""Architecture description of segofa (mirrors the presets models/segofa/segofa.py:351-467 and
the flags of unify_transformer.py:115-313 that shape the hot path) and the integer bucket
tables the model registers as buffers."""
import math
from dataclasses import dataclass, field
from typing import Tuple

import torch


@dataclass
class SegOFAConfig:
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
    code_image_size: int = 128
    decoder_input_type: str = "encoder_output"
    padding_idx: int = 1
    vocab: int = 59458  # len(dictionary) - num_seg (unify_transformer.py:402)

    @property
    def head_dim(self) -> int:
        return self.embed_dim // self.heads

    @property
    def attn_scaling(self) -> float:  # unify_multihead_attention.py:58
        return float(self.head_dim * self.attn_scale_factor) ** -0.5

    @property
    def pos_scaling(self) -> float:  # encoder_module.py:211
        return float(self.embed_dim / self.heads * self.attn_scale_factor) ** -0.5


ARCH_PRESETS = {
    # arch name: (embed_dim, ffn_dim, heads, enc_layers, dec_layers, resnet blocks, resnet_type)
    "segofa_tiny": (256, 1024, 4, 4, 4, (3, 4, 6), "resnet50"),
    "segofa_medium": (512, 2048, 8, 4, 4, (3, 4, 23), "resnet101"),
    "segofa_base": (768, 3072, 12, 6, 6, (3, 4, 23), "resnet101"),
    "segofa_large": (1024, 4096, 16, 12, 12, (3, 8, 36), "resnet152"),
    "segofa_huge": (1280, 5120, 16, 24, 12, (3, 8, 36), "resnet152"),
}
RESNET_BLOCKS = {"resnet50": (3, 4, 6), "resnet101": (3, 4, 23), "resnet152": (3, 8, 36)}


def preset(arch: str, **kw) -> SegOFAConfig:
    d, f, h, el, dl, rb, _ = ARCH_PRESETS[arch]
    cfg = dict(embed_dim=d, ffn_dim=f, heads=h, enc_layers=el, dec_layers=dl, resnet_blocks=rb)
    cfg.update(kw)
    return SegOFAConfig(**cfg)


def token_bucket_position(bucket_size: int, max_position: int = 1024) -> torch.Tensor:
    """Log-bucketed 1-D relative position index [max_position, max_position] (int64).
    Same arithmetic as encoder_module.py:71-84, evaluated once per distance d = i - j."""
    mid = bucket_size // 2
    d = torch.arange(-(max_position - 1), max_position, dtype=torch.long)
    mag = d.abs()
    mag = torch.where(mag < mid, torch.full_like(mag, mid - 1), mag)
    logb = torch.ceil(torch.log(mag / mid) / math.log((max_position - 1) / mid) * (mid - 1)) + mid
    far = logb.int() * torch.sign(d)
    per_distance = torch.where(mag <= mid, d, far.long()) + bucket_size - 1
    i = torch.arange(max_position).unsqueeze(1)
    j = torch.arange(max_position).unsqueeze(0)
    return per_distance[(i - j) + (max_position - 1)]


def image_bucket_position(bucket_size: int, num_relative_distance: int) -> torch.Tensor:
    """2-D relative position index over a bucket_size^2 grid plus the bos slot 0
    ([bs*bs+1, bs*bs+1] int64); same values as encoder_module.py:87-104."""
    n = bucket_size * bucket_size
    pos = torch.arange(n)
    y, x = pos // bucket_size, pos % bucket_size
    dy = y.unsqueeze(1) - y.unsqueeze(0) + (bucket_size - 1)
    dx = x.unsqueeze(1) - x.unsqueeze(0) + (bucket_size - 1)
    idx = torch.empty(n + 1, n + 1, dtype=torch.long)
    idx[1:, 1:] = dy * (2 * bucket_size - 1) + dx
    idx[0, :] = num_relative_distance - 3
    idx[:, 0] = num_relative_distance - 2
    idx[0, 0] = num_relative_distance - 1
    return idx
