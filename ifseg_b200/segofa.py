"""`segofa` model, B200-native.  Drop-in for models/segofa/segofa.py of alinlab/ifseg:

  * same registration: @register_model("segofa"), architectures segofa_{tiny,medium,base,large,huge}
    (segofa.py:25,351-467), staticmethod add_args exposing the flags of
    unify_transformer.py:115-313 + segofa.py:40-63, classmethod build_model(args, task)
    (unify_transformer.py:316-398);
  * same forward signature (segofa.py:69-88) and return value `(x, extra)`;
  * same parameter / buffer names (889 tensors for segofa_base), so checkpoints
    (`ofa_base.pt`, the reference's own) load and `state_dict()` round-trips;
  * same attributes other code reaches into (encoder.embed_tokens_bag, *.seg_embed_tokens,
    decoder.seg_projection, decoder.tie_seg_projection, encoder.dictionary, encoder.padding_idx).

The nn.Modules below are *parameter containers only*: none of their forward() methods is on
the hot path.  SegOFAModel.forward hands the tensors to ifseg_b200.engine.SegOFAEngine, which
issues a flat sequence of C-ABI kernel launches (libsegofa_b200.so).  There is no PyTorch
fallback: on a machine without the built extension or without a CUDA device forward() raises.
"""
import argparse
import logging
from typing import Optional

import torch
import torch.nn as nn

from .config import ARCH_PRESETS, RESNET_BLOCKS, SegOFAConfig, image_bucket_position, token_bucket_position
from .fairseq_compat import (
    FairseqEncoder,
    FairseqEncoderDecoderModel,
    FairseqIncrementalDecoder,
    StubDictionary,
    register_model,
    register_model_architecture,
)

logger = logging.getLogger(__name__)

_ACTS = ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"]

# (flags, kwargs) -- table-driven restatement of the reference CLI surface
_FLAGS = [
    (("--activation-fn",), dict(choices=_ACTS)),
    (("--dropout",), dict(type=float, metavar="D")),
    (("--attention-dropout",), dict(type=float, metavar="D")),
    (("--activation-dropout", "--relu-dropout"), dict(type=float, metavar="D")),
    (("--encoder-embed-path",), dict(type=str, metavar="STR")),
    (("--encoder-embed-dim",), dict(type=int, metavar="N")),
    (("--encoder-ffn-embed-dim",), dict(type=int, metavar="N")),
    (("--encoder-layers",), dict(type=int, metavar="N")),
    (("--encoder-attention-heads",), dict(type=int, metavar="N")),
    (("--encoder-normalize-before",), dict(action="store_true")),
    (("--encoder-learned-pos",), dict(action="store_true")),
    (("--bitfit",), dict(action="store_true", default=False)),
    (("--adapter",), dict(action="store_true")),
    (("--adapter-dim",), dict(type=int, metavar="N")),
    (("--encoder-prompt",), dict(action="store_true")),
    (("--encoder-prompt-type",), dict(type=str, choices=["prefix"])),
    (("--encoder-prompt-projection",), dict(action="store_true")),
    (("--encoder-prompt-length",), dict(type=int, metavar="N")),
    (("--encoder-prompt-dim",), dict(type=int, metavar="N")),
    (("--decoder-embed-path",), dict(type=str, metavar="STR")),
    (("--decoder-embed-dim",), dict(type=int, metavar="N")),
    (("--decoder-ffn-embed-dim",), dict(type=int, metavar="N")),
    (("--decoder-layers",), dict(type=int, metavar="N")),
    (("--decoder-attention-heads",), dict(type=int, metavar="N")),
    (("--decoder-learned-pos",), dict(action="store_true")),
    (("--decoder-normalize-before",), dict(action="store_true")),
    (("--decoder-output-dim",), dict(type=int, metavar="N")),
    (("--freeze-decoder",), dict(action="store_true")),
    (("--decoder-prompt",), dict(action="store_true")),
    (("--decoder-prompt-type",), dict(type=str, choices=["prefix"])),
    (("--decoder-prompt-length",), dict(type=int, metavar="N")),
    (("--decoder-prompt-projection",), dict(action="store_true")),
    (("--decoder-prompt-dim",), dict(type=int, metavar="N")),
    (("--share-decoder-input-output-embed",), dict(action="store_true")),
    (("--share-all-embeddings",), dict(action="store_true")),
    (("--no-token-positional-embeddings",), dict(action="store_true", default=False)),
    (("--adaptive-softmax-cutoff",), dict(metavar="EXPR")),
    (("--adaptive-softmax-dropout",), dict(type=float, metavar="D")),
    (("--layernorm-embedding",), dict(action="store_true")),
    (("--no-scale-embedding",), dict(action="store_true")),
    (("--checkpoint-activations",), dict(action="store_true")),
    (("--offload-activations",), dict(action="store_true")),
    (("--no-cross-attention",), dict(action="store_true", default=False)),
    (("--cross-self-attention",), dict(action="store_true", default=False)),
    (("--encoder-layerdrop",), dict(type=float, metavar="D", default=0)),
    (("--decoder-layerdrop",), dict(type=float, metavar="D", default=0)),
    (("--encoder-layers-to-keep",), dict(default=None)),
    (("--decoder-layers-to-keep",), dict(default=None)),
    (("--quant-noise-pq",), dict(type=float, metavar="D", default=0)),
    (("--quant-noise-pq-block-size",), dict(type=int, metavar="D", default=8)),
    (("--quant-noise-scalar",), dict(type=float, metavar="D", default=0)),
    (("--min-params-to-wrap",), dict(type=int, metavar="D", default=int(1e8))),
    (("--resnet-drop-path-rate",), dict(type=float)),
    (("--encoder-drop-path-rate",), dict(type=float)),
    (("--decoder-drop-path-rate",), dict(type=float)),
    (("--token-bucket-size",), dict(type=int)),
    (("--image-bucket-size",), dict(type=int)),
    (("--attn-scale-factor",), dict(type=float)),
    (("--freeze-resnet",), dict(type=str, default="false")),
    (("--freeze-entire-resnet",), dict(type=str, default="false")),
    (("--freeze-encoder-embedding",), dict(type=str, default="false")),
    (("--freeze-decoder-embedding",), dict(type=str, default="false")),
    (("--freeze-seg-embedding",), dict(type=str, default="false")),
    (("--freeze-encoder-transformer",), dict(type=str, default="false")),
    (("--freeze-encoder-transformer-layers",), dict(type=int, default=0)),
    (("--add-type-embedding",), dict(action="store_true")),
    (("--interpolate-position",), dict(action="store_true")),
    (("--resnet-type",), dict(choices=["resnet50", "resnet101", "resnet152"])),
    (("--resnet-model-path",), dict(type=str, metavar="STR")),
    (("--code-image-size",), dict(type=int)),
    (("--patch-layernorm-embedding",), dict(action="store_true")),
    (("--code-layernorm-embedding",), dict(action="store_true")),
    (("--entangle-position-embedding",), dict(action="store_true")),
    (("--disable-entangle",), dict(action="store_true")),
    (("--sync-bn",), dict(action="store_true")),
    (("--scale-attn",), dict(action="store_true")),
    (("--scale-fc",), dict(action="store_true")),
    (("--scale-heads",), dict(action="store_true")),
    (("--scale-resids",), dict(action="store_true")),
    (("--num-seg-tokens",), dict(type=int, default=150)),
    (("--decoder-type",), dict(type=str, default="surrogate")),
    (("--tie-seg-projection",), dict(type=str, default="false")),
    (("--decoder-input-type",), dict(type=str, default="encoder_input")),
    (("--patch-image-size",), dict(type=int, default=512)),
    (("--orig-patch-image-size",), dict(type=int, default=512)),
    # models/segofa/segofa.py:40-63
    (("--pooler-dropout",), dict(type=float, metavar="D")),
    (("--pooler-classifier",), dict(type=str, choices=["mlp", "linear"])),
    (("--pooler-activation-fn",), dict(choices=_ACTS)),
    (("--spectral-norm-classification-head",), dict(action="store_true")),
]


def str_bool(x) -> bool:
    """'true'/'false' string flags (encoder_module.py:50-57)."""
    if isinstance(x, bool):
        return x
    v = str(x).lower()
    if v not in ("true", "false"):
        raise ValueError(f"Unable to recognize string bool input: {x}")
    return v == "true"


# ----------------------------------------------------------------------------------------
# parameter containers
# ----------------------------------------------------------------------------------------
class _Container(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError(
            f"{type(self).__name__} is a parameter container; the segofa hot path runs through "
            "ifseg_b200.engine (CUDA C ABI), not through nn.Module.forward"
        )


def _embedding(n, d, padding_idx=None):
    m = nn.Embedding(n, d, padding_idx=padding_idx)
    nn.init.normal_(m.weight, mean=0.0, std=0.02)
    if padding_idx is not None:
        nn.init.constant_(m.weight[padding_idx], 0)
    return m


def _linear(i, o, bias=True):
    m = nn.Linear(i, o, bias)
    nn.init.normal_(m.weight, mean=0.0, std=0.02)
    if bias:
        nn.init.constant_(m.bias, 0.0)
    return m


class FrozenBatchNorm2d(_Container):
    """Buffers only (frozen_bn.py:29-34): weight, bias, running_mean, running_var; eps 1e-5."""

    def __init__(self, n, eps=1e-5):
        super().__init__()
        self.num_features, self.eps = n, eps
        self.register_buffer("weight", torch.ones(n))
        self.register_buffer("bias", torch.zeros(n))
        self.register_buffer("running_mean", torch.zeros(n))
        self.register_buffer("running_var", torch.ones(n) - eps)

    def _load_from_state_dict(self, state_dict, prefix, *a, **k):
        state_dict.pop(prefix + "num_batches_tracked", None)
        super()._load_from_state_dict(state_dict, prefix, *a, **k)


def _conv(i, o, k, stride=1, pad=0):
    m = nn.Conv2d(i, o, k, stride=stride, padding=pad, bias=False)
    nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
    return m


class Bottleneck(_Container):
    def __init__(self, inplanes, planes, stride, downsample):
        super().__init__()
        self.conv1, self.bn1 = _conv(inplanes, planes, 1), FrozenBatchNorm2d(planes)
        self.conv2, self.bn2 = _conv(planes, planes, 3, stride, 1), FrozenBatchNorm2d(planes)
        self.conv3, self.bn3 = _conv(planes, planes * 4, 1), FrozenBatchNorm2d(planes * 4)
        self.stride = stride
        self.downsample = (
            nn.Sequential(_conv(inplanes, planes * 4, 1, stride), FrozenBatchNorm2d(planes * 4)) if downsample else None
        )


class ResNetStem(_Container):
    """ResNet truncated after layer3 (stride 16, 1024 channels): resnet.py:140-229."""

    def __init__(self, blocks):
        super().__init__()
        self.conv1, self.bn1 = _conv(3, 64, 7, 2, 3), FrozenBatchNorm2d(64)
        inplanes = 64
        for li, (planes, n) in enumerate(zip((64, 128, 256), blocks)):
            layer = []
            for bi in range(n):
                layer.append(Bottleneck(inplanes, planes, 2 if (li > 0 and bi == 0) else 1, bi == 0))
                inplanes = planes * 4
            setattr(self, f"layer{li + 1}", nn.Sequential(*layer))


class MultiheadAttention(_Container):
    def __init__(self, d, heads, scale_heads):
        super().__init__()
        self.embed_dim, self.num_heads, self.head_dim = d, heads, d // heads
        self.c_attn = nn.Parameter(torch.ones(heads)) if scale_heads else None
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (_linear(d, d) for _ in range(4))


class EncoderLayer(_Container):
    def __init__(self, d, f, heads, scale_heads):
        super().__init__()
        self.self_attn = MultiheadAttention(d, heads, scale_heads)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.fc1, self.fc2 = _linear(d, f), _linear(f, d)
        self.attn_ln = nn.LayerNorm(d)
        self.ffn_layernorm = nn.LayerNorm(f)
        self.final_layer_norm = nn.LayerNorm(d)


class DecoderLayer(_Container):
    def __init__(self, d, f, heads, scale_heads):
        super().__init__()
        self.self_attn = MultiheadAttention(d, heads, scale_heads)
        self.self_attn_ln = nn.LayerNorm(d)
        self.cross_attn_ln = nn.LayerNorm(d)
        self.self_attn_layer_norm = nn.LayerNorm(d)
        self.encoder_attn = MultiheadAttention(d, heads, scale_heads)
        self.encoder_attn_layer_norm = nn.LayerNorm(d)
        self.ffn_layernorm = nn.LayerNorm(f)
        self.fc1, self.fc2 = _linear(d, f), _linear(f, d)
        self.final_layer_norm = nn.LayerNorm(d)
        self.need_attn = True


def _rel_tables(n, num, heads):
    return nn.ModuleList([_embedding(num, heads) for _ in range(n)])


class SegOFAEncoder(FairseqEncoder):
    """Parameter container mirroring encoder_module.py:106-262."""

    def __init__(self, args, cfg: SegOFAConfig, dictionary, embed_tokens, seg_embed_tokens):
        super().__init__(dictionary)
        self.args = args
        D, H = cfg.embed_dim, cfg.heads
        self.register_buffer("version", torch.Tensor([3]))
        self.padding_idx = embed_tokens.padding_idx
        self.max_source_positions = cfg.max_source_positions
        self.num_attention_heads = H
        self.embed_tokens = embed_tokens
        self.seg_embed_tokens = seg_embed_tokens
        self.embed_tokens_bag = nn.EmbeddingBag.from_pretrained(embed_tokens.weight, freeze=False)
        self.embed_tokens_bag.weight = embed_tokens.weight
        self.layernorm_embedding = nn.LayerNorm(D)
        self.type_embedding = _embedding(2, D)
        self.embed_images = ResNetStem(cfg.resnet_blocks)
        self.image_proj = _linear(1024, D)
        self.patch_layernorm_embedding = nn.LayerNorm(D)
        self.embed_positions = _embedding(cfg.max_source_positions + 2, D)
        self.embed_image_positions = _embedding(cfg.image_bucket_size ** 2 + 1, D)
        self.pos_ln, self.image_pos_ln = nn.LayerNorm(D), nn.LayerNorm(D)
        self.pos_q_linear, self.pos_k_linear = _linear(D, D), _linear(D, D)
        self.layers = nn.ModuleList([EncoderLayer(D, cfg.ffn_dim, H, True) for _ in range(cfg.enc_layers)])
        self.num_layers = cfg.enc_layers
        self.layer_norm = nn.LayerNorm(D)
        img_rel = (2 * cfg.image_bucket_size - 1) ** 2 + 3
        self.token_rel_pos_table_list = _rel_tables(cfg.enc_layers, 2 * cfg.token_bucket_size - 1, H)
        self.image_rel_pos_table_list = _rel_tables(cfg.enc_layers, img_rel, H)
        self.patch_image_size = cfg.patch_image_size
        self.orig_patch_image_size = cfg.orig_patch_image_size
        self.register_buffer("token_rp_bucket", token_bucket_position(cfg.token_bucket_size))
        self.register_buffer("image_rp_bucket", image_bucket_position(cfg.image_bucket_size, img_rel))
        # freezes of the shipped recipes (encoder_module.py:164-197): ResNet + image_proj
        if str_bool(getattr(args, "freeze_entire_resnet", "false")):
            for p in list(self.embed_images.parameters()) + list(self.image_proj.parameters()):
                p.requires_grad_(False)

    def max_positions(self):
        return self.max_source_positions


class SegOFADecoder(FairseqIncrementalDecoder):
    """Parameter container mirroring decoder_module.py:101-274 (surrogate decoder)."""

    def __init__(self, args, cfg: SegOFAConfig, dictionary, embed_tokens, seg_embed_tokens):
        super().__init__(dictionary)
        self.args = args
        D, H = cfg.embed_dim, cfg.heads
        self.register_buffer("version", torch.Tensor([3]))
        self.decoder_input_type = cfg.decoder_input_type
        self.decoder_type = getattr(args, "decoder_type", "surrogate")
        self.tie_seg_projection = str_bool(getattr(args, "tie_seg_projection", "false"))
        self.seg_embed_tokens = seg_embed_tokens
        self.seg_projection = _linear(D, cfg.num_seg, bias=False)
        if self.tie_seg_projection:
            self.seg_projection.weight = self.seg_embed_tokens.weight
        self.embed_dim = D
        self.padding_idx = embed_tokens.padding_idx
        self.max_target_positions = cfg.max_source_positions
        self.embed_tokens = embed_tokens
        self.layernorm_embedding = nn.LayerNorm(D)
        self.seg_bucket_size = cfg.patch_image_size // 16
        sb2 = self.seg_bucket_size ** 2 + 1
        self.embed_positions = _embedding(cfg.max_source_positions + 2, D)
        self.embed_image_positions = _embedding(cfg.image_bucket_size ** 2 + 1, D)
        self.embed_seg_positions = _embedding(sb2, D)
        self.pos_ln, self.image_pos_ln, self.seg_pos_ln = nn.LayerNorm(D), nn.LayerNorm(D), nn.LayerNorm(D)
        self.self_pos_q_linear, self.self_pos_k_linear = _linear(D, D), _linear(D, D)
        self.cross_pos_q_linear, self.cross_pos_k_linear = _linear(D, D), _linear(D, D)
        self.code_layernorm_embedding = nn.LayerNorm(D)
        self.layers = nn.ModuleList([DecoderLayer(D, cfg.ffn_dim, H, True) for _ in range(cfg.dec_layers)])
        self.num_layers = cfg.dec_layers
        self.layer_norm = nn.LayerNorm(D)
        img_rel = (2 * cfg.image_bucket_size - 1) ** 2 + 3
        seg_rel = (2 * self.seg_bucket_size - 1) ** 2 + 3
        self.token_rel_pos_table_list = _rel_tables(cfg.dec_layers, 2 * cfg.token_bucket_size - 1, H)
        self.image_rel_pos_table_list = _rel_tables(cfg.dec_layers, img_rel, H)
        self.seg_rel_pos_table_list = _rel_tables(cfg.dec_layers, seg_rel, H)
        ws = cfg.code_image_size // 8
        ipi = torch.arange(ws).unsqueeze(0).expand(ws, ws) + torch.arange(ws).unsqueeze(1) * cfg.image_bucket_size + 1
        ipi = torch.cat([torch.tensor([0]), ipi.reshape(-1), torch.tensor([1024] * 769)])
        self.register_buffer("seg_rp_bucket", image_bucket_position(self.seg_bucket_size, seg_rel))
        self.register_buffer("token_rp_bucket", token_bucket_position(cfg.token_bucket_size))
        self.register_buffer("image_rp_bucket", image_bucket_position(cfg.image_bucket_size, img_rel))
        self.register_buffer("image_position_idx", ipi)
        self.register_buffer("bin_id_offset", torch.tensor([dictionary.index("<bin_0>")]))
        self.register_buffer("seg_id_offset", torch.tensor([dictionary.index("<seg_0>")]))
        self.register_buffer("region_prefix", torch.tensor([976, 35]))

    def max_positions(self):
        return self.max_target_positions

    def output_projection(self, features):  # decoder_module.py:290-294 (host-level convenience)
        from . import ops

        out = ops.gemm(features.reshape(-1, features.shape[-1]).to(torch.bfloat16).contiguous(),
                       self.seg_projection.weight.to(torch.bfloat16).contiguous(), out_dtype=torch.float32)
        return out.view(*features.shape[:-1], -1)


# ----------------------------------------------------------------------------------------
# the model
# ----------------------------------------------------------------------------------------
_REQUIRED_TRUE = ("encoder_normalize_before", "decoder_normalize_before", "layernorm_embedding",
                  "patch_layernorm_embedding", "add_type_embedding", "scale_attn", "scale_fc", "scale_heads",
                  "no_scale_embedding", "share_all_embeddings")
_REQUIRED_FALSE = ("scale_resids", "adapter", "encoder_prompt", "decoder_prompt", "cross_self_attention",
                   "no_cross_attention", "sync_bn")


@register_model("segofa")
class SegOFAModel(FairseqEncoderDecoderModel):
    def __init__(self, args, encoder, decoder):
        super().__init__(encoder, decoder)
        self.args = args
        self.cfg: SegOFAConfig = encoder_cfg_from_args(args)
        self.classification_heads = nn.ModuleDict()
        if hasattr(self.encoder, "dictionary"):
            self.eos = self.encoder.dictionary.eos()
        self._engine = None
        self._train_engine = None

    # -- CLI / construction ---------------------------------------------------------------
    @staticmethod
    def add_args(parser):
        for flags, kw in _FLAGS:
            parser.add_argument(*flags, **kw)

    @classmethod
    def build_model(cls, args, task):
        """unify_transformer.py:316-398 for the configuration space the hot path supports."""
        base_architecture(args)
        for name in _REQUIRED_TRUE:
            if not getattr(args, name, False):
                raise NotImplementedError(
                    f"segofa_b200 implements the shipped IFSeg recipe; it requires --{name.replace('_', '-')}"
                )
        for name in _REQUIRED_FALSE:
            if getattr(args, name, False):
                raise NotImplementedError(f"segofa_b200 does not implement --{name.replace('_', '-')}")
        if getattr(args, "entangle_position_embedding", False) and not getattr(args, "disable_entangle", False):
            raise NotImplementedError("entangled position embeddings are not implemented (scripts pass --disable-entangle)")
        if getattr(args, "decoder_type", "surrogate") != "surrogate":
            raise NotImplementedError("only --decoder-type=surrogate exists (decoder_module.py:465-468)")
        if getattr(args, "activation_fn", "gelu") != "gelu":
            raise NotImplementedError("only --activation-fn=gelu is implemented")
        if getattr(args, "max_source_positions", None) is None:
            args.max_source_positions = 1024
        if getattr(args, "max_target_positions", None) is None:
            args.max_target_positions = 1024
        src_dict, tgt_dict = task.source_dictionary, task.target_dictionary
        if src_dict is not tgt_dict and src_dict != tgt_dict:
            raise ValueError("--share-all-embeddings requires a joined dictionary")
        cfg = encoder_cfg_from_args(args)
        cfg.vocab = len(src_dict) - args.num_seg_tokens
        cfg.padding_idx = src_dict.pad()
        args.vocab_size = cfg.vocab
        args.share_decoder_input_output_embed = True
        embed_tokens = _embedding(cfg.vocab, cfg.embed_dim, cfg.padding_idx)
        seg_embed_tokens = _embedding(cfg.num_seg, cfg.embed_dim)
        if str_bool(getattr(args, "freeze_encoder_embedding", "false")) or str_bool(
                getattr(args, "freeze_decoder_embedding", "false")):
            embed_tokens.weight.requires_grad = False
        if str_bool(getattr(args, "freeze_seg_embedding", "false")):
            seg_embed_tokens.weight.requires_grad = False
        encoder = SegOFAEncoder(args, cfg, src_dict, embed_tokens, seg_embed_tokens)
        decoder = SegOFADecoder(args, cfg, tgt_dict, embed_tokens, seg_embed_tokens)
        model = cls(args, encoder, decoder)
        model.cfg = cfg
        return model

    @classmethod
    def from_config(cls, arch="segofa_base", num_seg=15, image_size=480, **overrides):
        """Convenience constructor with the shipped flags (run_scripts/IFSeg/*.sh) -- used by
        bench.py / tests where no fairseq task exists."""
        args = shipped_args(arch, num_seg, image_size, **overrides)
        d = StubDictionary(num_seg)
        task = argparse.Namespace(source_dictionary=d, target_dictionary=d)
        return cls.build_model(args, task)

    # -- engine management ------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            from .engine import SegOFAEngine

            self._engine = SegOFAEngine(self)
        return self._engine

    def invalidate_engine(self):
        """Drop device-side derived weights (bf16 copies, folded BN, fused QKV).  Called on
        anything that may change parameters."""
        self._engine = None

    def parameters_changed(self):
        """Parameters were written in place by something other than the training engine's own optimizer step
        (load_state_dict, SegCriterion.lazy_initialization, an external optimizer on frozen tensors): drop the
        inference engine's derived copies and make the training engine re-derive EVERY operand, frozen ones included."""
        self.invalidate_engine()
        if self._train_engine is not None:
            self._train_engine.parameters_changed()

    def train_engine(self):
        """The training engine (flat fp32 master arena + hand-written backward); created on first use."""
        if self._train_engine is None:
            from .train_engine import SegOFATrainEngine

            self._train_engine = SegOFATrainEngine(self)
        return self._train_engine

    def train(self, mode: bool = True):
        self.invalidate_engine()
        return super().train(mode)

    def _apply(self, fn, *a, **k):
        self.invalidate_engine()
        self._train_engine = None  # the arena views are replaced by .to()/.cuda()/.float()
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, model_cfg=None, args=None, **kw):
        self.upgrade_state_dict_named(state_dict, "")
        out = super().load_state_dict(state_dict, strict, **kw)
        self.parameters_changed()  # values are copied in place (arena views stay valid); every derived copy is stale
        return out

    def upgrade_state_dict_named(self, state_dict, name):
        """Checkpoint compatibility (segofa.py:197-299, encoder_module.py:943-987,
        decoder_module.py:892-940): resize the token embedding to the current dictionary, drop
        mismatching seg tables, and fill every tensor the checkpoint lacks (attn_ln, seg tables,
        c_attn, ... absent from ofa_base.pt) with the freshly initialised one."""
        prefix = name + "." if name else ""
        own = self.state_dict()
        for k in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "encoder.embed_tokens_bag.weight"):
            t = state_dict.get(prefix + k)
            if t is not None and t.shape[0] != own[k].shape[0]:
                n_new = own[k].shape[0]
                if t.shape[0] > n_new:
                    state_dict[prefix + k] = t[:n_new]
                else:
                    extra = own[k][t.shape[0]:].to(t)
                    state_dict[prefix + k] = torch.cat([t, extra], dim=0)
        for k in ("encoder.seg_embed_tokens.weight", "decoder.seg_embed_tokens.weight", "decoder.seg_projection.weight",
                  "decoder.embed_seg_positions.weight", "decoder.seg_rp_bucket"):
            t = state_dict.get(prefix + k)
            if t is not None and t.shape != own[k].shape:
                logger.info("dropping %s from checkpoint (shape %s != %s)", k, tuple(t.shape), tuple(own[k].shape))
                del state_dict[prefix + k]
        for k in list(state_dict.keys()):
            if k.startswith(prefix + "decoder.seg_rel_pos_table_list.") and k[len(prefix):] in own and \
                    state_dict[k].shape != own[k[len(prefix):]].shape:
                del state_dict[k]
        for k, v in own.items():
            if prefix + k not in state_dict:
                state_dict[prefix + k] = v
        for k in [k for k in state_dict if k.startswith(prefix) and k[len(prefix):] not in own
                  and "classification_heads" not in k]:
            if k.endswith("num_batches_tracked") or ".output_projection." in k:
                del state_dict[k]

    # -- forward ---------------------------------------------------------------------------
    def forward(
        self,
        src_tokens: Optional[torch.Tensor] = None,
        src_lengths: Optional[torch.Tensor] = None,
        prev_output_tokens: Optional[torch.Tensor] = None,
        patch_images: Optional[torch.Tensor] = None,
        patch_images_2: Optional[torch.Tensor] = None,
        patch_masks: Optional[torch.Tensor] = None,
        code_masks: Optional[torch.Tensor] = None,
        sample_patch_num: Optional[int] = None,
        features_only: bool = False,
        full_context_alignment: bool = False,
        classification_head_name: Optional[str] = None,
        token_embeddings: Optional[torch.Tensor] = None,
        return_all_hiddens: bool = False,
        alignment_layer: Optional[int] = None,
        alignment_heads: Optional[int] = None,
        encoder_only: bool = False,
        aux_input: Optional[dict] = None,
    ):
        """Same signature and return contract as segofa.py:69-153."""
        for nm, val in (("patch_images_2", patch_images_2), ("code_masks", code_masks),
                        ("sample_patch_num", sample_patch_num), ("classification_head_name", classification_head_name),
                        ("token_embeddings", token_embeddings), ("alignment_heads", alignment_heads)):
            if val is not None:
                raise NotImplementedError(f"segofa_b200: `{nm}` is not on the IFSeg hot path (SURVEY.md s8)")
        if return_all_hiddens:
            raise NotImplementedError("segofa_b200: return_all_hiddens is not implemented")
        grad_mode = torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters())
        if grad_mode and src_tokens is not None and aux_input is not None:
            raise NotImplementedError("segofa_b200: one forward trains ONE branch (seg_criterion.py:178-192 never asks for "
                                      "gradients of both); call the other one under torch.no_grad()")
        # while a training engine exists its live-operand inference engine follows every optimizer update
        def no_grad_engine():
            return self._train_engine.live_inference_engine() if (self._train_engine is not None and self.training) else self.engine()

        x, extra = None, {}
        if src_tokens is not None and grad_mode:
            # supervised real-image branch (seg_criterion.py:188-192): the frozen ResNet + image_proj feed the same
            # hand-written forward/adjoint chain as the image-free branch (train_engine.py)
            if encoder_only or features_only:
                raise NotImplementedError("segofa_b200: encoder_only / features_only with gradients is not implemented")
            te = self.train_engine()
            x = te.real_logits(dict(src_tokens=src_tokens, patch_images=patch_images, patch_masks=patch_masks,
                                    prev_output_tokens=prev_output_tokens), causal=not full_context_alignment)
            B = src_tokens.shape[0]
            extra = {"attn": [None], "inner_states": [],
                     "encoder_returns": {"image_embed_shape": [te.last_grid], "encoder_states": [], "src_tokens": [],
                                         "src_lengths": [], "patch_images": [], "encoder_embedding": []}}
        elif src_tokens is not None:
            eng = no_grad_engine()
            enc = eng.encode(src_tokens, patch_images=patch_images, patch_masks=patch_masks)
            if encoder_only:
                return eng.encoder_out_dict(enc)
            x, extra = eng.decode(enc, prev_output_tokens, full_context_alignment=full_context_alignment,
                                  features_only=features_only)
            extra["encoder_returns"] = eng.encoder_out_dict(enc)
        if aux_input is not None and grad_mode:
            # image-free branch with gradients: forward keeps the activations, backward is the hand-written adjoint
            # chain (train_engine.py); always causal (full_context_alignment is not forwarded, segofa.py:145-149)
            ax = self.train_engine().imfree_logits(aux_input)
            extra["aux_output"] = (ax, {"attn": [None], "inner_states": []})
        elif aux_input is not None:
            eng = no_grad_engine()
            aenc = eng.encode(aux_input.get("src_tokens"), bag_tokens=aux_input.get("patch_images"),
                              bag_offsets=aux_input.get("patch_masks"))
            ax, aextra = eng.decode(aenc, aux_input.get("prev_output_tokens"), full_context_alignment=False)
            extra["aux_output"] = (ax, aextra)
        return x, extra


def encoder_cfg_from_args(args) -> SegOFAConfig:
    return SegOFAConfig(
        embed_dim=args.encoder_embed_dim,
        ffn_dim=args.encoder_ffn_embed_dim,
        heads=args.encoder_attention_heads,
        enc_layers=args.encoder_layers,
        dec_layers=args.decoder_layers,
        resnet_blocks=RESNET_BLOCKS[args.resnet_type],
        num_seg=args.num_seg_tokens,
        patch_image_size=args.patch_image_size,
        orig_patch_image_size=args.orig_patch_image_size,
        token_bucket_size=args.token_bucket_size,
        image_bucket_size=args.image_bucket_size,
        attn_scale_factor=args.attn_scale_factor,
        max_source_positions=getattr(args, "max_source_positions", 1024) or 1024,
        code_image_size=args.code_image_size,
        decoder_input_type=getattr(args, "decoder_input_type", "encoder_input"),
    )


# ----------------------------------------------------------------------------------------
# architecture presets: fill every attribute that is still unset (segofa.py:351-467)
# ----------------------------------------------------------------------------------------
_LARGE_DEFAULTS = dict(
    encoder_embed_path=None, encoder_normalize_before=True, encoder_learned_pos=True, decoder_embed_path=None,
    decoder_normalize_before=True, decoder_learned_pos=True, attention_dropout=0.0, relu_dropout=0.0, dropout=0.0,
    max_target_positions=1024, max_source_positions=1024, adaptive_softmax_cutoff=None, adaptive_softmax_dropout=0,
    share_decoder_input_output_embed=True, share_all_embeddings=True, no_scale_embedding=True,
    layernorm_embedding=True, activation_fn="gelu", pooler_activation_fn="tanh", pooler_dropout=0.0,
    pooler_classifier="mlp", resnet_drop_path_rate=0.0, encoder_drop_path_rate=0.0, decoder_drop_path_rate=0.0,
    token_bucket_size=256, image_bucket_size=42, freeze_encoder_embedding=False, freeze_decoder_embedding=False,
    add_type_embedding=True, attn_scale_factor=2, code_image_size=128, patch_layernorm_embedding=True,
    code_layernorm_embedding=True, entangle_position_embedding=False, disable_entangle=False, sync_bn=False,
    scale_attn=False, scale_fc=False, scale_heads=False, scale_resids=False, orig_patch_image_size=256,
)


def _setdefaults(args, **kw):
    for k, v in kw.items():
        if getattr(args, k, None) is None:
            setattr(args, k, v)


def _apply_preset(args, arch):
    d, f, h, el, dl, _, rtype = ARCH_PRESETS[arch]
    _setdefaults(args, encoder_embed_dim=d, encoder_ffn_embed_dim=f, encoder_layers=el, encoder_attention_heads=h,
                 decoder_layers=dl, decoder_attention_heads=h, resnet_type=rtype)
    _setdefaults(args, decoder_embed_dim=args.encoder_embed_dim, decoder_ffn_embed_dim=args.encoder_ffn_embed_dim)
    _setdefaults(args, decoder_output_dim=args.decoder_embed_dim, decoder_input_dim=args.decoder_embed_dim)
    _setdefaults(args, **_LARGE_DEFAULTS)


def base_architecture(args):
    """Residual defaults of unify_transformer.py:495-559 that the presets leave unset."""
    _setdefaults(args, activation_dropout=0.0, no_token_positional_embeddings=False, adaptive_input=False,
                 no_cross_attention=False, cross_self_attention=False, encoder_prompt=False, decoder_prompt=False,
                 adapter=False, tie_adaptive_weights=False, checkpoint_activations=False, offload_activations=False,
                 encoder_layers_to_keep=None, decoder_layers_to_keep=None, encoder_layerdrop=0, decoder_layerdrop=0,
                 quant_noise_pq=0, quant_noise_pq_block_size=8, quant_noise_scalar=0)


@register_model_architecture("segofa", "segofa_large")
def segofa_large_architecture(args):
    _apply_preset(args, "segofa_large")


@register_model_architecture("segofa", "segofa_base")
def segofa_base_architecture(args):
    _apply_preset(args, "segofa_base")


@register_model_architecture("segofa", "segofa_huge")
def segofa_huge_architecture(args):
    _apply_preset(args, "segofa_huge")


@register_model_architecture("segofa", "segofa_medium")
def segofa_medium_architecture(args):
    _apply_preset(args, "segofa_medium")


@register_model_architecture("segofa", "segofa_tiny")
def segofa_tiny_architecture(args):
    _apply_preset(args, "segofa_tiny")


_ARCH_FNS = {
    "segofa_tiny": segofa_tiny_architecture, "segofa_medium": segofa_medium_architecture,
    "segofa_base": segofa_base_architecture, "segofa_large": segofa_large_architecture,
    "segofa_huge": segofa_huge_architecture,
}

SHIPPED_FLAGS = dict(  # run_scripts/IFSeg/coco_unseen.sh:76-136
    encoder_normalize_before=True, decoder_normalize_before=True, share_all_embeddings=True,
    share_decoder_input_output_embed=True, layernorm_embedding=True, patch_layernorm_embedding=True,
    code_layernorm_embedding=True, add_type_embedding=True, scale_attn=True, scale_fc=True, scale_heads=True,
    disable_entangle=True, dropout=0.1, attention_dropout=0.0, encoder_drop_path_rate=0.1,
    decoder_drop_path_rate=0.1, resnet_drop_path_rate=0.0, freeze_encoder_embedding="true",
    freeze_decoder_embedding="true", freeze_seg_embedding="true", freeze_entire_resnet="true",
    tie_seg_projection="true", decoder_type="surrogate", decoder_input_type="encoder_output",
)


def shipped_args(arch, num_seg, image_size, **overrides):
    """Namespace as train.py hands it to build_model for run_scripts/IFSeg/*.sh."""
    parser = argparse.ArgumentParser(argument_default=argparse.SUPPRESS)
    SegOFAModel.add_args(parser)
    args = parser.parse_args([])
    for k in [k for k, v in vars(args).items() if v is None]:
        delattr(args, k)
    for k, v in SHIPPED_FLAGS.items():
        setattr(args, k, v)
    args.num_seg_tokens = num_seg
    args.patch_image_size = image_size
    args.orig_patch_image_size = image_size
    args.arch = arch
    for k, v in overrides.items():
        setattr(args, k, v)
    _ARCH_FNS[arch](args)
    return args
