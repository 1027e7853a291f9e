"""CPU: host-side mirror of the reference interface (registration, flags, names, checkpoints)."""
import argparse
import json
import os

import pytest
import torch

from helpers import GOLD


@pytest.fixture(scope="module")
def model():
    from ifseg_b200.segofa import SegOFAModel

    return SegOFAModel.from_config("segofa_base", 15, 128)


def test_state_dict_matches_reference_manifest(model):
    man = json.load(open(os.path.join(GOLD, "manifest_base_c15_s128.json")))
    sd = model.state_dict()
    assert [m[0] for m in man] == list(sd.keys())  # same names, same order (889 tensors)
    for name, shape, dtype in man:
        assert list(sd[name].shape) == shape, name
        assert str(sd[name].dtype).replace("torch.", "") == dtype, name
    assert len(man) == 889


def test_manifest_150_classes():
    from ifseg_b200.segofa import SegOFAModel

    m = SegOFAModel.from_config("segofa_base", 150, 64)
    man = json.load(open(os.path.join(GOLD, "manifest_base_c150_s64_b2.json")))
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == {n: s for n, s, _ in man}


@pytest.mark.parametrize("arch,name", [("segofa_tiny", "tiny_c15_s64"), ("segofa_medium", "medium_c15_s64"),
                                       ("segofa_large", "large_c15_s64"), ("segofa_huge", "huge_c15_s64")])
def test_other_architectures_match_reference_manifest(arch, name):
    """state-dict names, order, shapes and dtypes of the reference's tiny / large models (oracle/make_golden.py)."""
    from ifseg_b200.segofa import SegOFAModel

    m = SegOFAModel.from_config(arch, 15, 64)
    man = json.load(open(os.path.join(GOLD, f"manifest_{name}.json")))
    sd = m.state_dict()
    assert [x[0] for x in man] == list(sd.keys())
    for n, shape, dtype in man:
        assert list(sd[n].shape) == shape and str(sd[n].dtype).replace("torch.", "") == dtype, n


def test_parameter_counts_and_freezes(model):
    total = sum(p.numel() for p in model.parameters())
    trainable = sum(p.numel() for p in model.parameters() if p.requires_grad)
    assert total == 182_235_880
    assert trainable == 108_320_808  # SURVEY.md s8a (108.3 M with the shipped freezes)
    assert model.encoder.embed_tokens.weight is model.decoder.embed_tokens.weight
    assert model.encoder.embed_tokens_bag.weight is model.encoder.embed_tokens.weight
    assert model.decoder.seg_projection.weight is model.decoder.seg_embed_tokens.weight
    assert model.decoder.tie_seg_projection is True


def test_registration_and_cli_surface():
    from ifseg_b200 import fairseq_compat as fc
    from ifseg_b200.segofa import SegOFAModel

    assert fc.MODEL_REGISTRY["segofa"] is SegOFAModel
    for arch in ("segofa_tiny", "segofa_medium", "segofa_base", "segofa_large", "segofa_huge"):
        assert fc.ARCH_MODEL_REGISTRY[arch] is SegOFAModel and callable(fc.ARCH_CONFIG_REGISTRY[arch])
    parser = argparse.ArgumentParser(argument_default=argparse.SUPPRESS)
    SegOFAModel.add_args(parser)
    flags = [a for act in parser._actions for a in act.option_strings if a.startswith("--") and a != "--help"]
    assert len(flags) == 85 + 4 + 1  # unify_transformer.py flags + segofa.py:40-63 (+ --relu-dropout alias)
    # the flags run_scripts/IFSeg/coco_unseen.sh:76-136 passes to the model must all parse
    ns = parser.parse_args(
        "--encoder-normalize-before --decoder-normalize-before --share-decoder-input-output-embed "
        "--share-all-embeddings --layernorm-embedding --patch-layernorm-embedding --code-layernorm-embedding "
        "--resnet-drop-path-rate=0.0 --encoder-drop-path-rate=0.1 --decoder-drop-path-rate=0.1 --dropout=0.1 "
        "--attention-dropout=0.0 --add-type-embedding --scale-attn --scale-fc --scale-heads --disable-entangle "
        "--patch-image-size=512 --orig-patch-image-size=512 --tie-seg-projection=true --decoder-type=surrogate "
        "--decoder-input-type=encoder_output --num-seg-tokens=15 --freeze-encoder-embedding=true "
        "--freeze-decoder-embedding=true --freeze-seg-embedding=true --freeze-entire-resnet=true".split())
    assert ns.scale_heads and ns.num_seg_tokens == 15 and not hasattr(ns, "no_scale_embedding")


@pytest.mark.parametrize("arch,d,f,h,el,dl,rn", [
    ("segofa_tiny", 256, 1024, 4, 4, 4, "resnet50"), ("segofa_medium", 512, 2048, 8, 4, 4, "resnet101"),
    ("segofa_base", 768, 3072, 12, 6, 6, "resnet101"), ("segofa_large", 1024, 4096, 16, 12, 12, "resnet152"),
    ("segofa_huge", 1280, 5120, 16, 24, 12, "resnet152")])
def test_arch_presets(arch, d, f, h, el, dl, rn):
    from ifseg_b200 import fairseq_compat as fc

    ns = argparse.Namespace()
    fc.ARCH_CONFIG_REGISTRY[arch](ns)
    assert (ns.encoder_embed_dim, ns.encoder_ffn_embed_dim, ns.encoder_attention_heads, ns.encoder_layers,
            ns.decoder_layers, ns.resnet_type) == (d, f, h, el, dl, rn)
    assert ns.no_scale_embedding and ns.token_bucket_size == 256 and ns.image_bucket_size == 42
    assert ns.attn_scale_factor == 2 and ns.orig_patch_image_size == 256  # segofa.py:419 default
    ns2 = argparse.Namespace(encoder_layers=3, dropout=0.3)
    fc.ARCH_CONFIG_REGISTRY[arch](ns2)
    assert ns2.encoder_layers == 3 and ns2.dropout == 0.3  # explicit values win over the preset


def test_bucket_tables_match_oracle():
    from ifseg_b200.config import image_bucket_position, token_bucket_position
    from oracle import restated as R

    assert torch.equal(token_bucket_position(256), R.make_token_bucket_position(256))
    for bs in (4, 30, 42):
        n = (2 * bs - 1) ** 2 + 3
        assert torch.equal(image_bucket_position(bs, n), R.make_image_bucket_position(bs, n))


def test_forward_fails_loudly_without_cuda(model):
    """No CPU fallback: the product path must raise, not silently compute on the host."""
    from ifseg_b200.synthetic import synthetic_inputs

    model.eval()
    inp = synthetic_inputs(model.cfg, 1, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        with torch.no_grad():
            model(**inp)
    model.train()
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # supervised branch: the training engine is CUDA-only
        model(**inp)
    with pytest.raises(RuntimeError, match="no CPU fallback"):  # the training engine is CUDA-only as well
        model(aux_input=dict(src_tokens=inp["src_tokens"], patch_images=inp["src_tokens"], patch_masks=inp["src_lengths"],
                             prev_output_tokens=inp["prev_output_tokens"]))
    model.eval()
    with pytest.raises(NotImplementedError):
        with torch.no_grad():
            model(**inp, sample_patch_num=10)


def test_ops_refuse_cpu_tensors():
    from ifseg_b200 import ops

    a = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.gemm(a, a)


def test_checkpoint_upgrade_fills_missing_and_resizes(model):
    """ofa_base.pt-style checkpoint: 59457 embedding rows, no attn_ln / seg tables / c_attn."""
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    for k in list(sd):
        if ".attn_ln." in k or "seg_" in k or k.endswith("c_attn") or ".ffn_layernorm." in k:
            del sd[k]
    for k in ("encoder.embed_tokens.weight", "decoder.embed_tokens.weight", "encoder.embed_tokens_bag.weight"):
        sd[k] = sd[k][:59457]
    sd["encoder.embed_images.bn1.num_batches_tracked"] = torch.tensor(0)
    res = model.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert model.encoder.embed_tokens.weight.shape[0] == 59458
    # a checkpoint trained with another class count: seg tables are dropped, not loaded
    sd2 = {k: v.clone() for k, v in model.state_dict().items()}
    sd2["decoder.seg_projection.weight"] = torch.zeros(150, 768)
    sd2["encoder.seg_embed_tokens.weight"] = torch.zeros(150, 768)
    sd2["decoder.seg_embed_tokens.weight"] = torch.zeros(150, 768)
    model.load_state_dict(sd2, strict=True)
    assert model.decoder.seg_projection.weight.shape == (15, 768)


def test_unsupported_flags_are_rejected():
    from ifseg_b200.segofa import SegOFAModel

    with pytest.raises(NotImplementedError):
        SegOFAModel.from_config("segofa_base", 15, 128, scale_heads=False)
    with pytest.raises(NotImplementedError):
        SegOFAModel.from_config("segofa_base", 15, 128, adapter=True)
