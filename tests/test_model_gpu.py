"""-m gpu: the CUDA segofa forward (through SegOFAModel.forward -> C ABI) against
  (a) the committed golden fixtures = outputs of the UNMODIFIED reference model (fp32), and
  (b) the travelling oracle (oracle/restated.py) on the same seeded weights/inputs.

Tolerance.  north_star asks for <=1e-3 rel on bf16 logits; SURVEY.md s7 shows that figure is
only meaningful between runs with identical rounding points -- the reference's OWN bf16 run
differs from its fp32 run by 1.4-1.6e-2 rel-L2 (stored per fixture as ref_bf16_rel_l2).  The
gate used here: rel-L2(ours, reference fp32) <= 0.75 x that reference bf16 noise floor, i.e. we
must be strictly closer to the fp32 truth than the reference's bf16 path is (measured values are
printed; they are ~3-5e-3 because the residual stream, softmax, LayerNorm and GELU run in
fp32).  Masks: given identical logits the upsample+argmax kernel must be BIT-EXACT against the
reference mask; end-to-end, every pixel whose oracle top-2 margin exceeds 4x the max logit error
must agree.
"""
import pytest
import torch

from helpers import build_cuda_model, load_golden, oracle_cfg, rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def case15(cuda_device):
    g = load_golden("base_c15_s128")
    model, sd = build_cuda_model(g["arch"], g["num_seg"], g["image_size"], g["weight_seed"])
    return g, model, sd


def _inputs(g, model, tokens=None):
    from ifseg_b200.synthetic import synthetic_inputs

    inp = synthetic_inputs(model.cfg, g["batch"], g["image_size"], g["input_seed"], g["src_tokens"][0])
    if tokens is not None:
        inp["src_tokens"] = tokens
    return {k: v.cuda() for k, v in inp.items()}


def test_logits_vs_reference_golden(case15):
    g, model, _ = case15
    with torch.no_grad():
        x, extra = model(**_inputs(g, model))
    assert x.shape == g["logits"].shape and x.dtype == torch.float32
    err = rel_l2(x, g["logits"])
    print(f"rel-L2 vs reference fp32: {err:.3e}; reference bf16 noise floor {g['ref_bf16_rel_l2']:.3e}")
    assert err <= 0.75 * g["ref_bf16_rel_l2"]
    enc = extra["encoder_returns"]
    assert enc["image_embed_shape"][0] == (8, 8)
    assert rel_l2(enc["encoder_out"][0].transpose(0, 1), g["encoder_out"]) <= 0.75 * g["ref_bf16_rel_l2"]
    e_feat = rel_l2(enc["image_embed_before_proj"][0], g["resnet_features"])
    e_enc = rel_l2(enc["encoder_out"][0].transpose(0, 1), g["encoder_out"])
    print(f"stage errors: resnet features {e_feat:.3e}, encoder_out {e_enc:.3e}")
    assert e_feat <= 2e-2


def test_full_context_and_padding_variants(case15):
    g, model, _ = case15
    with torch.no_grad():
        x_fc, _ = model(**_inputs(g, model), full_context_alignment=True)
        toks = g["src_tokens"].clone()
        toks[0, -5:] = 1
        x_pad, _ = model(**_inputs(g, model, toks))
    assert rel_l2(x_fc, g["logits_full_context"]) <= 0.75 * g["ref_bf16_rel_l2"]
    assert rel_l2(x_pad, g["logits_padded"]) <= 0.75 * g["ref_bf16_rel_l2"]
    assert rel_l2(x_fc, g["logits"]) > 1e-3  # the causal mask matters


def test_mask_parity(case15):
    g, model, _ = case15
    S, hp = g["image_size"], g["image_size"] // 16
    with torch.no_grad():
        x, _ = model(**_inputs(g, model))
        eng = model.engine()
        # (1) same logits in -> bit-identical mask out
        m_ref_logits = eng.predict_mask(g["logits"].cuda(), (hp, hp), (S, S))
        assert torch.equal(m_ref_logits.cpu().to(torch.int16), g["mask"].view(-1, S, S))
        # (2) end to end: disagreement only where the reference's own margin is within the logit error
        m = eng.predict_mask(x, (hp, hp), (S, S)).cpu()
    from oracle import restated as R

    up = R.upsample_logits(g["logits"], hp, hp, S, S)[:, :-1]
    top2 = up.topk(2, dim=-1).values
    margin = (top2[..., 0] - top2[..., 1]).view(-1, S, S)
    max_err = (x.cpu() - g["logits"]).abs().max().item()
    disagree = m != g["mask"].view(-1, S, S).long()
    print(f"mask agreement {1 - disagree.float().mean().item():.5f}, max logit err {max_err:.3e}, "
          f"largest margin among disagreeing pixels {margin[disagree].max().item() if disagree.any() else 0:.3e}")
    assert not (disagree & (margin > 4 * max_err)).any()
    assert disagree.float().mean().item() < 0.05


def test_second_config_batch2_150_classes(cuda_device):
    g = load_golden("base_c150_s64_b2")
    model, sd = build_cuda_model(g["arch"], g["num_seg"], g["image_size"], g["weight_seed"])
    with torch.no_grad():
        x, _ = model(**_inputs(g, model))
    err = rel_l2(x, g["logits"])
    print(f"rel-L2 vs reference fp32: {err:.3e}; floor {g['ref_bf16_rel_l2']:.3e}")
    assert err <= 0.75 * g["ref_bf16_rel_l2"]


def test_against_travelling_oracle_other_seed(cuda_device):
    """fresh weights/inputs not covered by a fixture: oracle (CPU fp32) vs CUDA, 96x96, B=2."""
    from oracle import restated as R
    from ifseg_b200.synthetic import synthetic_inputs

    model, sd = build_cuda_model("segofa_base", 15, 96, seed=3)
    inp = synthetic_inputs(model.cfg, 2, 96, seed=5)
    with torch.no_grad():
        ref, _ = R.segofa_forward(sd, oracle_cfg(model.cfg), inp["src_tokens"], inp["patch_images"], inp["patch_masks"])
        x, _ = model(**{k: v.cuda() for k, v in inp.items()})
    err = rel_l2(x, ref)
    print(f"rel-L2 vs oracle: {err:.3e}")
    assert err <= 1.2e-2


def test_launches_are_ours(case15):
    from ifseg_b200 import ops

    g, model, _ = case15
    ops.reset_launch_count()
    with torch.no_grad():
        model(**_inputs(g, model))
    n = ops.launch_count()
    print("kernel launches per forward:", n)
    assert n > 100


def test_image_free_branch_vs_reference_golden(case15):
    """models/segofa/segofa.py:136-151: aux_input -> extra['aux_output'][0] (always causal)."""
    g, model, _ = case15
    aux = {k: v.cuda() for k, v in g["aux_input"].items()}
    with torch.no_grad():
        x, extra = model(aux_input=aux)
    assert x is None
    ax = extra["aux_output"][0]
    err = rel_l2(ax, g["aux_logits"])
    print(f"image-free branch rel-L2 vs reference fp32: {err:.3e}")
    assert ax.shape == g["aux_logits"].shape and err <= 0.75 * g["ref_bf16_rel_l2"]


def test_criterion_eval_branch(case15):
    """SegCriterion.forward (eval): areas == compute_metric on the oracle-upsampled logits of OUR logits,
    display CE == F.cross_entropy; keys as seg_criterion.py:222-233."""
    import types
    from ifseg_b200.fairseq_compat import StubDictionary
    from ifseg_b200.seg_criterion import SegCriterion, derive_metrics
    from oracle import restated as R

    g, model, _ = case15
    S, C = g["image_size"], g["num_seg"]
    task = types.SimpleNamespace(target_dictionary=StubDictionary(C),
                                 cfg=types.SimpleNamespace(num_seg_tokens=C, category_list=",".join(f"c{i}" for i in range(C))))
    crit = SegCriterion(task, init_seg_with_text="false")
    gen = torch.Generator().manual_seed(9)
    classes = torch.randint(0, C + 1, (g["batch"], S * S), generator=gen)  # includes the 'unknown' id C
    target = torch.cat([classes + 59457, torch.full((g["batch"], 1), 2)], 1)
    inp = _inputs(g, model)
    sample = {"net_input": inp, "target": target, "ntokens": 1, "nsentences": g["batch"]}
    loss, sample_size, log = crit(model, sample)
    for k in ("loss", "imfree_loss", "seg_loss", "ntokens", "nsentences", "sample_size", "area_intersect",
              "area_pred_label", "area_label", "area_union", "nll_loss"):
        assert k in log
    with torch.no_grad():
        x, _ = model(**inp)
    up = R.upsample_logits(x.cpu(), S // 16, S // 16, S, S)[:, :-1].reshape(-1, C)
    t = classes.reshape(-1)
    valid = t < C
    ai, ap, al, au = R.compute_metric(up[valid], t[valid])
    assert torch.equal(log["area_intersect"].cpu(), ai) and torch.equal(log["area_pred_label"].cpu(), ap)
    assert torch.equal(log["area_label"].cpu(), al) and torch.equal(log["area_union"].cpu(), au)
    ref_loss = torch.nn.functional.cross_entropy(up[valid], t[valid])
    assert abs(loss.item() - ref_loss.item()) < 1e-3
    m = derive_metrics(ai, ap, al, au)
    assert 0 <= m["mIoU"] <= 1 and 0 <= m["aAcc"] <= 1
    # validation post-processing (resnet_iters > 0): the propagated probabilities give a second set of areas
    crit_pp = SegCriterion(task, resnet_iters=5, resnet_topk=3, init_seg_with_text="false")
    _, _, log_pp = crit_pp(model, sample)
    for k in ("area_intersect_resnet_postprocess", "area_pred_label_resnet_postprocess", "area_label_resnet_postprocess",
              "area_union_resnet_postprocess"):
        assert k in log_pp and log_pp[k].shape == (C,)
    assert torch.equal(log_pp["area_label_resnet_postprocess"], log_pp["area_label"])
    assert log_pp["area_pred_label_resnet_postprocess"].sum() == log_pp["area_pred_label"].sum()
    # image-free loss value (training forward) against the oracle's imfree_loss on OUR aux logits
    aux = {k: v.cuda() for k, v in g["aux_input"].items()}
    t2s = torch.cat([torch.randint(0, C + 1, (g["batch"], S * S), generator=gen) + 59457,
                     torch.full((g["batch"], 1), 2)], 1)
    val = crit.imfree_loss_value(model, {"aux_input": aux, "text2seg_target": t2s})
    with torch.no_grad():
        _, extra = model(aux_input=aux)
    ocfg = R.SegOFAConfig(num_seg=C, patch_image_size=S)
    ref_val = R.imfree_loss(extra["aux_output"][0].cpu(), t2s, ocfg)
    assert abs(val.item() - ref_val.item()) < 1e-3


@pytest.mark.parametrize("arch,num_seg,size,batch", [("segofa_large", 150, 96, 2), ("segofa_tiny", 15, 64, 3),
                                                     ("segofa_medium", 171, 80, 1)])
def test_other_architectures_vs_oracle(cuda_device, arch, num_seg, size, batch):
    """Every registered architecture (ResNet-50/101/152 stems, D in {256,512,1024}, 4..16 heads) against the
    CPU oracle on fresh seeded weights."""
    from oracle import restated as R
    from ifseg_b200.synthetic import synthetic_inputs

    model, sd = build_cuda_model(arch, num_seg, size, seed=11)
    inp = synthetic_inputs(model.cfg, batch, size, seed=13)
    torch.set_num_threads(8)
    with torch.no_grad():
        ref, _ = R.segofa_forward(sd, oracle_cfg(model.cfg), inp["src_tokens"], inp["patch_images"], inp["patch_masks"])
        x, _ = model(**{k: v.cuda() for k, v in inp.items()})
    err = rel_l2(x, ref)
    print(f"{arch}: rel-L2 vs oracle {err:.3e}")
    assert x.shape == ref.shape and err <= 1.5e-2


def test_serving_session_matches_eager(case15):
    """SegmentationSession (CUDA graph + pinned pipelined I/O) returns the same masks as the eager API."""
    from ifseg_b200.serving import SegmentationSession
    from ifseg_b200.synthetic import synthetic_inputs

    g, model, _ = case15
    S = g["image_size"]
    sess = SegmentationSession(model, 2, S, g["src_tokens"][0])
    batches = [synthetic_inputs(model.cfg, 2, S, seed=100 + i)["patch_images"] for i in range(5)]
    outs = [m.clone() for m in sess.infer_stream(batches)]
    assert len(outs) == 5
    eng = model.engine()
    for imgs, m in zip(batches, outs):
        with torch.no_grad():
            x, extra = model(src_tokens=g["src_tokens"][:1].repeat(2, 1).cuda(), patch_images=imgs.cuda(),
                             patch_masks=torch.ones(2, dtype=torch.bool).cuda(),
                             prev_output_tokens=torch.zeros(2, 1, dtype=torch.long).cuda())
            ref = eng.predict_mask(x, extra["encoder_returns"]["image_embed_shape"][0], (S, S))
        assert torch.equal(m, ref.cpu())
    assert torch.equal(sess.infer(batches[0]), outs[0])
    # a yielded tensor stays intact across the NEXT yield (three rotating pinned buffers): hold it without cloning
    held = []
    for k, m in enumerate(sess.infer_stream(batches)):
        if held:
            torch.cuda.synchronize()  # everything the generator has queued so far has landed
            assert torch.equal(held[-1], outs[k - 1]), "result buffer overwritten while the caller still held it"
        held.append(m)


@pytest.mark.parametrize("Hi,Wi", [(128, 192), (96, 160), (160, 128)])
def test_general_patch_grids_vs_oracle(cuda_device, Hi, Wi):
    """Validation keeps the image aspect ratio: patch grids that differ from the orig / seg_bucket
    grid go through the interpolated position tables (encoder_module.py:358-370, 782-808; decoder_module.py:541-550,
    601-625).  8x12 (more patches than the 8x8 orig grid), 6x10 (fewer, other shape), 10x8."""
    from oracle import restated as R
    from ifseg_b200.synthetic import synthetic_inputs

    model, sd = build_cuda_model("segofa_base", 15, 128, seed=3)
    inp = synthetic_inputs(model.cfg, 2, 128, seed=5)
    g = torch.Generator().manual_seed(Hi + Wi)
    images = torch.randn(2, 3, Hi, Wi, generator=g)
    torch.set_num_threads(8)
    with torch.no_grad():
        ref, _ = R.segofa_forward(sd, oracle_cfg(model.cfg), inp["src_tokens"], images, inp["patch_masks"])
        x, extra = model(src_tokens=inp["src_tokens"].cuda(), src_lengths=inp["src_lengths"].cuda(),
                         prev_output_tokens=inp["prev_output_tokens"].cuda(), patch_images=images.cuda(),
                         patch_masks=inp["patch_masks"].cuda())
    hp, wp = extra["encoder_returns"]["image_embed_shape"][0]
    assert (hp, wp) == (Hi // 16, Wi // 16) and x.shape == ref.shape == (2, hp * wp + 1, 15)
    err = rel_l2(x, ref)
    assert err < 1.2e-2, err
    mask = model.engine().predict_mask(x, (hp, wp), (Hi, Wi))
    assert torch.equal(mask.cpu().view(2, -1), R.predict_mask(x.cpu(), hp, wp, Hi, Wi))
