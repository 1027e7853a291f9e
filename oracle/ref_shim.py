"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (ifseg_b200/).

Loads the UNMODIFIED reference model (`/root/reference/models/segofa/*.py`) in a
container that has no fairseq install, by providing an in-memory stub `fairseq`
package.  Recipe follows SURVEY.md Appendix A:

  * vendored leaf modules are loaded *by path, unchanged* from
    /root/reference/custom_fairseq/fairseq/... (gelu, fairseq_dropout, layer_norm,
    layer_drop, grad_multiply, quant_noise, incremental_decoding_utils,
    models/fairseq_{encoder,decoder,incremental_decoder});
  * eight tiny symbols are restated here (each cites the reference line it follows).

This file only works where /root/reference exists (the build container).  It is
used by oracle/make_golden.py to (a) validate oracle/restated.py and (b) emit the
golden fixtures under tests/golden/.  Nothing on the GPU box imports it.
"""
import argparse
import importlib.util
import os
import sys
import types

import torch
import torch.nn as nn
import torch.nn.functional as F

REF_ROOT = os.environ.get("IFSEG_REFERENCE_ROOT", "/root/reference")
_CF = os.path.join(REF_ROOT, "custom_fairseq", "fairseq")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "models", "segofa"))


def _load_by_path(modname, relpath):
    spec = importlib.util.spec_from_file_location(modname, os.path.join(_CF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


class StubDictionary:
    """Sizes/special ids of the task dictionary after SegmentationTask.setup_task
    (tasks/mm_tasks/segmentation.py:109-136): 50260 dict.txt rows + 4 specials
    + <mask> + 8192 <code_i> + 1000 <bin_i> = 59457, then num_seg+1 <seg_i>."""

    def __init__(self, num_seg):
        self.num_seg = num_seg

    def __len__(self):
        return 59457 + self.num_seg + 1

    def __eq__(self, other):
        return self is other

    def __contains__(self, sym):
        return True

    def bos(self):
        return 0

    def pad(self):
        return 1

    def eos(self):
        return 2

    def unk(self):
        return 3

    def index(self, sym):
        if sym == "<bin_0>":
            return 58457
        if sym == "<seg_0>":
            return 59457
        raise KeyError(sym)


def install_stub_fairseq():
    if "fairseq" in sys.modules and getattr(sys.modules["fairseq"], "_ifseg_stub", False):
        return
    fs = types.ModuleType("fairseq")
    fs.__path__ = []
    fs._ifseg_stub = True
    sys.modules["fairseq"] = fs

    # ---- fairseq.utils: restated helpers -------------------------------------------------
    utils = types.ModuleType("fairseq.utils")

    def softmax(x, dim, onnx_trace=False):  # custom_fairseq/fairseq/utils.py:510-514
        return F.softmax(x, dim=dim, dtype=torch.float32)

    def fill_with_neg_inf(t):  # utils.py:398-400
        return t.float().fill_(float("-inf")).type_as(t)

    def new_arange(x, *size):  # utils.py:690-697
        if len(size) == 0:
            size = x.size()
        return torch.arange(size[-1], device=x.device).expand(*size).contiguous()

    def get_available_activation_fns():  # utils.py:563-571
        return ["relu", "gelu", "gelu_fast", "gelu_accurate", "tanh", "linear"]

    def item(t):
        return t.item() if hasattr(t, "item") else t

    utils.softmax = softmax
    utils.fill_with_neg_inf = fill_with_neg_inf
    utils.new_arange = new_arange
    utils.get_available_activation_fns = get_available_activation_fns
    utils.item = item
    sys.modules["fairseq.utils"] = utils
    fs.utils = utils

    # ---- fairseq.modules: vendored leaf files, loaded unchanged ---------------------------
    modules = types.ModuleType("fairseq.modules")
    modules.__path__ = []
    sys.modules["fairseq.modules"] = modules
    fs.modules = modules
    gelu_m = _load_by_path("fairseq.modules.gelu", "modules/gelu.py")
    qn = _load_by_path("fairseq.modules.quant_noise", "modules/quant_noise.py")
    fd = _load_by_path("fairseq.modules.fairseq_dropout", "modules/fairseq_dropout.py")
    ln = _load_by_path("fairseq.modules.layer_norm", "modules/layer_norm.py")
    ld = _load_by_path("fairseq.modules.layer_drop", "modules/layer_drop.py")
    gm = _load_by_path("fairseq.modules.grad_multiply", "modules/grad_multiply.py")
    modules.gelu = gelu_m.gelu
    modules.gelu_accurate = gelu_m.gelu_accurate
    modules.FairseqDropout = fd.FairseqDropout
    modules.LayerNorm = ln.LayerNorm
    modules.LayerDropModuleList = ld.LayerDropModuleList
    modules.GradMultiply = gm.GradMultiply
    modules.quant_noise = qn

    class AdaptiveSoftmax(nn.Module):  # imported, unused on the surrogate path
        pass

    class BaseLayer(nn.Module):
        pass

    class SinusoidalPositionalEmbedding(nn.Module):
        pass

    modules.AdaptiveSoftmax = AdaptiveSoftmax
    modules.BaseLayer = BaseLayer
    modules.SinusoidalPositionalEmbedding = SinusoidalPositionalEmbedding

    ca = types.ModuleType("fairseq.modules.checkpoint_activations")
    ca.checkpoint_wrapper = lambda m, **kw: m
    sys.modules["fairseq.modules.checkpoint_activations"] = ca

    tse = types.ModuleType("fairseq.modules.transformer_sentence_encoder")

    def init_bert_params(module):  # modules/transformer_sentence_encoder.py:21-53
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.bias is not None:
                module.bias.data.zero_()
        if isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()

    tse.init_bert_params = init_bert_params
    sys.modules["fairseq.modules.transformer_sentence_encoder"] = tse

    def get_activation_fn(activation):  # utils.py:540-560
        if activation == "relu":
            return F.relu
        if activation == "gelu":
            return gelu_m.gelu
        if activation in ("gelu_fast", "gelu_accurate"):
            return gelu_m.gelu_accurate
        if activation == "tanh":
            return torch.tanh
        if activation == "linear":
            return lambda x: x
        raise RuntimeError("--activation-fn {} not supported".format(activation))

    utils.get_activation_fn = get_activation_fn

    # ---- fairseq.distributed --------------------------------------------------------------
    dist = types.ModuleType("fairseq.distributed")
    dist.fsdp_wrap = lambda m, **kw: m
    sys.modules["fairseq.distributed"] = dist
    fs.distributed = dist

    # ---- fairseq.incremental_decoding_utils / fairseq.models -----------------------------
    _load_by_path("fairseq.incremental_decoding_utils", "incremental_decoding_utils.py")
    models = types.ModuleType("fairseq.models")
    models.__path__ = []
    sys.modules["fairseq.models"] = models
    fs.models = models
    enc = _load_by_path("fairseq.models.fairseq_encoder", "models/fairseq_encoder.py")
    dec = _load_by_path("fairseq.models.fairseq_decoder", "models/fairseq_decoder.py")
    models.FairseqEncoder = enc.FairseqEncoder
    models.FairseqDecoder = dec.FairseqDecoder
    inc = _load_by_path(
        "fairseq.models.fairseq_incremental_decoder", "models/fairseq_incremental_decoder.py"
    )
    models.FairseqIncrementalDecoder = inc.FairseqIncrementalDecoder

    class BaseFairseqModel(nn.Module):  # models/fairseq_model.py:36-60 (minimal)
        @classmethod
        def add_args(cls, parser):
            pass

    class FairseqEncoderDecoderModel(BaseFairseqModel):  # fairseq_model.py:270-288
        def __init__(self, encoder, decoder):
            super().__init__()
            self.encoder = encoder
            self.decoder = decoder

    models.MODEL_REGISTRY = {}
    models.ARCH_CONFIG_REGISTRY = {}

    def register_model(name, dataclass=None):  # models/__init__.py:109-150
        def deco(cls):
            models.MODEL_REGISTRY[name] = cls
            return cls

        return deco

    def register_model_architecture(model_name, arch_name):  # models/__init__.py:153-194
        def deco(fn):
            models.ARCH_CONFIG_REGISTRY[arch_name] = fn
            return fn

        return deco

    models.BaseFairseqModel = BaseFairseqModel
    models.FairseqEncoderDecoderModel = FairseqEncoderDecoderModel
    models.register_model = register_model
    models.register_model_architecture = register_model_architecture


_SHIPPED_FLAGS = dict(  # run_scripts/IFSeg/coco_unseen.sh:76-136
    encoder_normalize_before=True,
    decoder_normalize_before=True,
    share_all_embeddings=True,
    share_decoder_input_output_embed=True,
    layernorm_embedding=True,
    patch_layernorm_embedding=True,
    code_layernorm_embedding=True,
    add_type_embedding=True,
    scale_attn=True,
    scale_fc=True,
    scale_heads=True,
    disable_entangle=True,
    dropout=0.1,
    attention_dropout=0.0,
    encoder_drop_path_rate=0.1,
    decoder_drop_path_rate=0.1,
    resnet_drop_path_rate=0.0,
    freeze_encoder_embedding="true",
    freeze_decoder_embedding="true",
    freeze_seg_embedding="true",
    freeze_entire_resnet="true",
    tie_seg_projection="true",
    decoder_type="surrogate",
    decoder_input_type="encoder_output",
)


def shipped_args(arch, num_seg, image_size, parser_cls_add_args):
    """argparse namespace as train.py would hand it to build_model."""
    # fairseq adds the model flags to a group created with argument_default=SUPPRESS
    # (custom_fairseq/fairseq/options.py:142-149): flags that are not passed and carry no explicit
    # default are ABSENT from the namespace, so the arch preset's getattr defaults apply
    # (e.g. no_scale_embedding=True -> embed_scale == 1).
    parser = argparse.ArgumentParser(argument_default=argparse.SUPPRESS)
    parser_cls_add_args(parser)
    args = parser.parse_args([])
    for k in [k for k, v in vars(args).items() if v is None]:
        delattr(args, k)
    for k, v in _SHIPPED_FLAGS.items():
        setattr(args, k, v)
    args.num_seg_tokens = num_seg
    args.patch_image_size = image_size
    args.orig_patch_image_size = image_size
    return args


def build_reference_model(arch="segofa_base", num_seg=15, image_size=128, seed=0, overrides=None):
    """Returns (model.eval() in fp32, args).  Random init exactly as the reference does it."""
    assert reference_available(), "reference tree not mounted"
    install_stub_fairseq()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    from models.segofa.segofa import SegOFAModel  # noqa: the unmodified reference
    import fairseq.models as fm

    args = shipped_args(arch, num_seg, image_size, SegOFAModel.add_args)
    for k, v in (overrides or {}).items():
        setattr(args, k, v)
    fm.ARCH_CONFIG_REGISTRY[arch](args)
    d = StubDictionary(num_seg)
    task = types.SimpleNamespace(source_dictionary=d, target_dictionary=d)
    torch.manual_seed(seed)
    model = SegOFAModel.build_model(args, task)
    model.eval()
    return model, args
