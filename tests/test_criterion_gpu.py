"""-m gpu: criterion behaviour that launches kernels -- automatic lazy seg-token initialisation (seg_criterion.py:173-176,
373-407), `_lowres` metrics (:273-281, 314-321), and parameter changes made AFTER the engines exist (lazy init,
load_state_dict) reaching the frozen operands of the training engine."""
import pytest
import torch

from helpers import build_cuda_model, load_prompts, make_task, rel_l2

pytestmark = pytest.mark.gpu


def _sample(model, B, S, C, seed=9, lowres=False):
    from ifseg_b200.synthetic import synthetic_inputs

    inp = {k: v.cuda() for k, v in synthetic_inputs(model.cfg, B, S, seed=1, src_tokens=load_prompts()["15"]).items()}
    gen = torch.Generator().manual_seed(seed)
    target = torch.cat([torch.randint(0, C + 1, (B, S * S), generator=gen) + 59457, torch.full((B, 1), 2)], 1)
    sample = {"net_input": inp, "target": target, "ntokens": 1, "nsentences": B}
    if lowres:
        P = (S // 16) ** 2
        low = torch.randint(0, C + 1, (B, P), generator=gen) + 59457
        low[0, :3] = 1  # padding
        sample["downsampled_target"] = torch.cat([low, torch.full((B, 1), 2)], 1)
    return sample


def test_first_forward_runs_lazy_initialisation_for_model_and_ema(cuda_device):
    from ifseg_b200.seg_criterion import SegCriterion

    C, S, B = 15, 64, 2
    model, sd = build_cuda_model("segofa_base", C, S, seed=0)
    ema, _ = build_cuda_model("segofa_base", C, S, seed=0)
    task = make_task(C)
    crit = SegCriterion(task)  # init_seg_with_text defaults to 'true'
    sample = _sample(model, B, S, C)
    with torch.no_grad():
        before, _ = model(**sample["net_input"])  # builds (and caches) the inference engine with the random classifier
    old = model.decoder.seg_projection.weight.detach().clone()
    crit(model, sample, update_num=7, ema_model=ema)
    assert crit.iter == 7 and crit.effective_iter == 7  # counter restored from update_num (:174), then advanced once
    ids = crit.class_name_token_ids()
    table = sd["encoder.embed_tokens.weight"]
    want = torch.stack([table[t].mean(0) for t in ids])
    for m in (model, ema):
        assert torch.allclose(m.encoder.seg_embed_tokens.weight.cpu(), want, atol=1e-6)
        assert m.decoder.seg_projection.weight is m.decoder.seg_embed_tokens.weight  # the tie survives (in-place write)
    assert not torch.equal(old, model.decoder.seg_projection.weight)
    with torch.no_grad():
        after, _ = model(**sample["net_input"])
    assert rel_l2(after, before) > 0.1  # the cached engine was dropped: the logits use the text-initialised classifier
    w = model.decoder.seg_projection.weight.detach().clone()
    crit(model, sample, update_num=8, ema_model=ema)  # runs once only
    assert crit.iter == 8 and torch.equal(w, model.decoder.seg_projection.weight)


def test_lowres_metric_family(cuda_device):
    from oracle import restated as R
    from ifseg_b200.seg_criterion import SegCriterion

    C, S, B = 15, 64, 2
    model, _ = build_cuda_model("segofa_base", C, S, seed=0)
    crit = SegCriterion(make_task(C), init_seg_with_text="false", upscale_lprobs="false")
    sample = _sample(model, B, S, C, lowres=True)
    _, _, log = crit(model, sample)
    with torch.no_grad():
        x, _ = model(**sample["net_input"])
    t = sample["downsampled_target"][:, :-1].reshape(-1)
    keep = (t != 1) & (t != 59457 + C)
    ai, ap, al, au = R.compute_metric(x[:, :-1].reshape(-1, C).cpu()[keep], (t - 59457)[keep])
    assert torch.equal(log["area_intersect_lowres"].cpu(), ai) and torch.equal(log["area_pred_label_lowres"].cpu(), ap)
    assert torch.equal(log["area_label_lowres"].cpu(), al) and torch.equal(log["area_union_lowres"].cpu(), au)
    # upscale_lprobs=false: the display loss is the CE of the patch-grid logits against the patch-grid labels (:342-343)
    ref = torch.nn.functional.cross_entropy(x[:, :-1].reshape(-1, C).cpu()[keep], (t - 59457)[keep])
    assert abs(log["nll_loss"].item() - ref.item()) < 1e-4


def test_parameter_changes_after_engine_construction_reach_frozen_operands(cuda_device):
    """the training engine copies frozen weights (e.g. the tied, frozen seg_projection) when it is built; a
    later lazy initialisation / load_state_dict must reach those copies, in training AND in the live no-grad engine."""
    from ifseg_b200 import ops
    from ifseg_b200.seg_criterion import SegCriterion, class_targets
    from ifseg_b200.synthetic import generate_state_dict

    C, S, B = 15, 64, 2
    model, sd = build_cuda_model("segofa_base", C, S, seed=0)
    model.train()
    eng = model.train_engine()  # built FIRST (SegOFATrainer.__init__ does the same)
    eng.stochastic = False
    names = torch.full((C, 4), 1, dtype=torch.long)
    lens = torch.ones(C, dtype=torch.int32)
    names[:, 0] = torch.arange(C) + 100
    prompt = torch.tensor(load_prompts()["15"], dtype=torch.long)
    bag, ends, target = ops.artificial_sample(names.cuda(), lens.cuda(), B, S // 16, S, seed=7)
    aux = dict(src_tokens=prompt.unsqueeze(0).repeat(B, 1).cuda(), patch_images=bag, patch_masks=ends,
               prev_output_tokens=torch.zeros(B, 1, dtype=torch.long).cuda())
    l0 = eng.forward_train(aux)["logits"].clone()
    crit = SegCriterion(make_task(C))
    crit.lazy_initialization(model, crit.class_name_token_ids())
    l1 = eng.forward_train(aux)["logits"].clone()
    assert rel_l2(l1, l0) > 0.1, "the training engine still uses the stale classifier"
    # the same model evaluated by a freshly built engine must agree with the long-lived one
    model.eval()
    with torch.no_grad():
        _, extra = model(aux_input=aux)
    assert rel_l2(l1, extra["aux_output"][0]) < 1.2e-2  # (train forward vs folded no-grad forward: bf16 noise level)
    # load_state_dict after construction: stem / classifier / trainable weights all follow
    model.train()
    sd2 = generate_state_dict(model.cfg, 5)
    model.load_state_dict(sd2, strict=True)
    l2 = eng.forward_train(aux)["logits"].clone()
    fresh, _ = build_cuda_model("segofa_base", C, S, seed=5)
    with torch.no_grad():
        _, ex2 = fresh(aux_input=aux)
    assert rel_l2(l2, ex2["aux_output"][0]) < 1.2e-2
    from ifseg_b200.synthetic import synthetic_inputs

    inp = {k: v.cuda() for k, v in synthetic_inputs(model.cfg, B, S, seed=1, src_tokens=load_prompts()["15"]).items()}
    with torch.no_grad():
        a, _ = model(**inp)       # model.training: goes through the training engine's live no-grad engine (stem included)
        b, _ = fresh(**inp)
    assert rel_l2(a, b) < 1.2e-2
