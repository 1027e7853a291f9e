"""-m gpu: the training step of the image-free branch (forward with saved activations + hand-written
backward, ifseg_b200/train_engine.py) against the oracle's autograd and the gradient fixture generated
from the unmodified reference (tests/golden/golden_grads_*.pt, oracle/make_golden.py:run_grad_case).

Tolerance (written here, as north_star asks): the reference's own bf16 run differs from its fp32 run by
4.8e-2 global gradient rel-L2 (per tensor: median 1.5e-2, max 1.4e-1; stored in the fixture).  Gate: our
global gradient rel-L2 against the fp32 oracle is below that noise floor, the loss agrees to 1e-2 relative."""
import pytest
import torch

from helpers import load_golden, oracle_cfg, rel_l2

pytestmark = pytest.mark.gpu


def _setup(name="base_c150_s64_b2"):
    import os

    from helpers import GOLD
    from ifseg_b200.seg_criterion import class_targets
    from ifseg_b200.segofa import SegOFAModel
    from ifseg_b200.synthetic import generate_state_dict
    from ifseg_b200.train_engine import SegOFATrainEngine

    g = load_golden(name)
    gg = torch.load(os.path.join(GOLD, f"golden_grads_{name}.pt"), map_location="cpu")
    C, S, B = g["num_seg"], g["image_size"], g["batch"]
    model = SegOFAModel.from_config(g["arch"], C, S)
    sd = generate_state_dict(model.cfg, 0)
    model.load_state_dict(sd, strict=True)
    model = model.cuda()
    eng = SegOFATrainEngine(model, stochastic=False)  # eval-mode numerics: the gradient-parity configuration
    t2s = gg["text2seg_target"].long()
    tgt = class_targets(t2s[:, :-1].reshape(B, S, S), 59457, C).cuda()
    aux = {k: v.cuda() for k, v in g["aux_input"].items()}
    return g, gg, model, sd, eng, aux, tgt, t2s


def test_imfree_train_step_gradients_vs_oracle(cuda_device):
    from oracle import restated as R

    g, gg, model, sd, eng, aux, tgt, t2s = _setup()
    loss, logits = eng.forward_backward(aux, tgt)
    torch.cuda.synchronize()
    assert abs(loss.item() - gg["loss"]) < 1e-2 * gg["loss"], (loss.item(), gg["loss"])
    assert rel_l2(logits, g["aux_logits"]) <= 0.75 * g["ref_bf16_rel_l2"]

    ours = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    # golden norms from the reference itself
    ratios = []
    for k, gr in ours.items():
        assert k in gg["grad_norms"], f"{k}: the reference produces no gradient here"
        n_ref = gg["grad_norms"][k]
        if n_ref > 1e-8:
            ratios.append((gr.norm().item() / n_ref, k))
    lo, hi = min(ratios), max(ratios)
    assert 0.8 < lo[0] and hi[0] < 1.25, (lo, hi)
    for k, ref in gg["small_grads"].items():
        if k in ours and gg["grad_norms"][k] > 1e-8:
            assert rel_l2(ours[k], ref) < 0.2, (k, rel_l2(ours[k], ref))
    # full comparison against the oracle's autograd (fp32, CPU)
    torch.set_num_threads(8)
    ocfg = oracle_cfg(model.cfg)
    sd_g = {k: (v.clone().requires_grad_() if k in ours else v) for k, v in sd.items()}
    x_or, _ = R.segofa_forward_aux(sd_g, ocfg, g["aux_input"])
    R.imfree_loss(x_or, t2s, ocfg).backward()
    num = den = 0.0
    worst = (0.0, None)
    per = []
    for k, gr in ours.items():
        go = sd_g[k].grad
        if go is None or go.norm().item() < 1e-8:
            assert gr.norm().item() < 1e-5, (k, gr.norm().item())  # k_proj biases: exactly zero in exact arithmetic
            continue
        num += ((gr - go) ** 2).sum().item()
        den += (go ** 2).sum().item()
        e = rel_l2(gr, go)
        per.append(e)
        if e > worst[0]:
            worst = (e, k)
    glob = (num / den) ** 0.5
    per.sort()
    print(f"\ngradient parity: {len(per)} tensors, global rel-L2 {glob:.3e} (reference bf16 floor "
          f"{gg['ref_bf16_grad_rel_l2']:.3e}), per-tensor median {per[len(per) // 2]:.3e}, worst {worst[0]:.3e} {worst[1]}")
    assert glob <= gg["ref_bf16_grad_rel_l2"], glob
    assert worst[0] < 0.25, worst
    # every tensor the reference gives a gradient to is either produced or on the documented bias-path list
    missing = [k for k in gg["grad_norms"] if k not in ours]
    assert not missing, missing
    assert len(ours) == len(gg["grad_norms"]) == 380  # exactly the tensors the reference gives a gradient to


def test_train_loop_decreases_loss_and_keeps_views(cuda_device):
    g, gg, model, sd, eng, aux, tgt, t2s = _setup()
    first = None
    for it in range(6):
        loss, _ = eng.forward_backward(aux, tgt)
        gn = eng.optimizer_step(lr=2e-4, weight_decay=0.01, clip_norm=1.0)
        if first is None:
            first = loss.item()
            assert abs(gn.item() - sum(v ** 2 for v in gg["grad_norms"].values()) ** 0.5) < 0.08 * gn.item()
    last, _ = eng.forward_backward(aux, tgt, backward=False)
    assert last.item() < first - 0.05, (first, last.item())
    # the nn.Parameters are views of the flat master buffer: state_dict sees the updated weights
    p = model.encoder.layers[0].fc1.weight
    assert p.data_ptr() >= eng.arena.flat32.data_ptr()
    assert not torch.equal(p.detach().cpu(), sd["encoder.layers.0.fc1.weight"])
    assert torch.isfinite(eng.arena.flat32).all()


def test_criterion_training_branch_through_autograd(cuda_device):
    """The drop-in call sequence of task.train_step (segmentation.py:190-222): model.train();
    loss, sample_size, log = criterion(model, sample); optimizer.backward(loss) -> loss.backward().
    Gradients arrive in param.grad (views of the arena) and equal the fused forward_backward path."""
    import types

    from ifseg_b200.fairseq_compat import StubDictionary
    from ifseg_b200.seg_criterion import SegCriterion
    from ifseg_b200.synthetic import synthetic_inputs

    g, gg, model, sd, eng0, aux, tgt, t2s = _setup()
    # reference gradients of the fused path (own engine instance of the same model)
    loss0, _ = eng0.forward_backward(aux, tgt)
    ref = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}
    model._train_engine = eng0
    model.train()
    C, S, B = g["num_seg"], g["image_size"], g["batch"]
    task = types.SimpleNamespace(target_dictionary=StubDictionary(C),
                                 cfg=types.SimpleNamespace(num_seg_tokens=C, category_list=",".join(f"c{i}" for i in range(C))))
    crit = SegCriterion(task, init_seg_with_text="false")
    inp = {k: v.cuda() for k, v in synthetic_inputs(model.cfg, B, S, seed=1, src_tokens=g["src_tokens"][0]).items()}
    gen = torch.Generator().manual_seed(9)
    target = torch.cat([torch.randint(0, C + 1, (B, S * S), generator=gen) + 59457, torch.full((B, 1), 2)], 1)
    sample = {"net_input": inp, "aux_input": aux, "target": target, "text2seg_target": t2s, "ntokens": 1, "nsentences": B}
    model.zero_grad(set_to_none=True)
    loss, sample_size, log = crit(model, sample)
    assert loss.requires_grad and sample_size == 1
    for k in ("loss", "imfree_loss", "seg_loss", "area_intersect", "area_union", "nll_loss"):
        assert k in log
    assert abs(loss.item() - loss0.item()) < 1e-5
    loss.backward()
    torch.cuda.synchronize()
    n = 0
    for k, p in model.named_parameters():
        if k in ref:
            assert p.grad is not None and p.grad.data_ptr() >= eng0.arena.grad32.data_ptr()
            assert rel_l2(p.grad, ref[k]) < 2e-3 or ref[k].norm().item() < 1e-6, (k, rel_l2(p.grad, ref[k]))
            n += 1
    assert n == len(ref) and n >= 323
    # autograd semantics: a second backward accumulates
    loss2, _, _ = crit(model, sample)
    loss2.backward()
    k = "encoder.layers.0.fc1.weight"
    assert rel_l2(dict(model.named_parameters())[k].grad, 2 * ref[k]) < 2e-3
    # eval after training uses the updated parameters through a fresh inference engine
    model.eval()
    with torch.no_grad():
        x, _ = model(**inp)
    assert torch.isfinite(x).all()


def test_stochastic_training_is_seeded_and_learns(cuda_device):
    """Recipe noise (dropout 0.1, DropPath up to 0.1): same seed -> identical gradients, other seed -> different,
    and the native training loop still reduces the (deterministic) loss."""
    from ifseg_b200.train_engine import SegOFATrainEngine

    g, gg, model, sd, eng, aux, tgt, t2s = _setup()
    assert eng.drop_p == pytest.approx(0.1) and eng.enc_dpr[-1] == pytest.approx(0.1) and eng.enc_dpr[0] == 0.0

    def grads(seed):
        e = SegOFATrainEngine(model, stochastic=True, seed=seed)
        loss, _ = e.forward_backward(aux, tgt)
        return loss.item(), e.arena.grad32.clone()

    l1, g1 = grads(5)
    l2, g2 = grads(5)
    l3, g3 = grads(6)
    # same masks -> same loss and gradients up to the summation order of the fp32 atomics
    assert abs(l1 - l2) < 1e-5 and rel_l2(g1, g2) < 1e-5
    assert rel_l2(g1, g3) > 1e-2
    assert abs(l1 - gg["loss"]) < 0.5  # noisy forward, same ballpark
    e = SegOFATrainEngine(model, stochastic=True, seed=3)
    e.stochastic = False
    first, _ = e.forward_backward(aux, tgt, backward=False)
    e.stochastic = True
    for _ in range(8):
        e.forward_backward(aux, tgt)
        e.optimizer_step(lr=2e-4, weight_decay=0.01, clip_norm=1.0)
    e.stochastic = False
    last, _ = e.forward_backward(aux, tgt, backward=False)
    assert last.item() < first.item() - 0.05, (first.item(), last.item())


def _supervised_sample(g, model, aux, t2s):
    from ifseg_b200.synthetic import synthetic_inputs

    C, S, B = g["num_seg"], g["image_size"], g["batch"]
    inp_host = synthetic_inputs(model.cfg, B, S, seed=1, src_tokens=g["src_tokens"][0])
    gen = torch.Generator().manual_seed(9)
    # dictionary ids at the network input resolution; class C = the ignore label (seg_criterion.py:296-298), eos last
    target = torch.cat([torch.randint(0, C + 1, (B, S * S), generator=gen) + 59457, torch.full((B, 1), 2)], 1)
    sample = {"net_input": {k: v.cuda() for k, v in inp_host.items()}, "aux_input": aux, "target": target,
              "text2seg_target": t2s, "ntokens": 1, "nsentences": B}
    return inp_host, target, sample


def test_supervised_real_image_branch_gradients_vs_oracle(cuda_device):
    """--unsupervised-segmentation=false (seg_criterion.py:188-192): the real-image loss carries the gradient.  ResNet
    and image_proj are frozen (--freeze-entire-resnet=true, encoder_module.py:191-197), every other tensor gets the
    same adjoint chain as the image-free branch.  Gate as for that branch: global gradient rel-L2 against the fp32
    oracle's autograd below the reference's own bf16 noise floor."""
    import types

    from oracle import restated as R

    from ifseg_b200.fairseq_compat import StubDictionary
    from ifseg_b200.seg_criterion import SegCriterion

    g, gg, model, sd, eng, aux, tgt, t2s = _setup()
    C, S, B = g["num_seg"], g["image_size"], g["batch"]
    inp_host, target, sample = _supervised_sample(g, model, aux, t2s)
    model._train_engine = eng
    model.train()
    task = types.SimpleNamespace(target_dictionary=StubDictionary(C),
                                 cfg=types.SimpleNamespace(num_seg_tokens=C, category_list=",".join(f"c{i}" for i in range(C))))
    crit = SegCriterion(task, init_seg_with_text="false", unsupervised_segmentation="false")
    model.zero_grad(set_to_none=True)
    loss, sample_size, log = crit(model, sample)
    assert loss.requires_grad and sample_size == 1 and float(log["imfree_loss"]) == 0.0
    for k in ("loss", "seg_loss", "area_intersect", "area_union", "nll_loss"):
        assert k in log
    loss.backward()
    torch.cuda.synchronize()
    ours = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    assert len(ours) == 380 and not any(k.startswith(("encoder.embed_images", "encoder.image_proj")) for k in ours)

    torch.set_num_threads(8)
    ocfg = oracle_cfg(model.cfg)
    sd_g = {k: (v.clone().requires_grad_() if k in ours else v) for k, v in sd.items()}
    x_or, _ = R.segofa_forward(sd_g, ocfg, inp_host["src_tokens"], inp_host["patch_images"], inp_host["patch_masks"],
                               inp_host["prev_output_tokens"])
    loss_or = R.imfree_loss(x_or, target, ocfg)
    loss_or.backward()
    assert abs(loss.item() - loss_or.item()) < 1e-2 * loss_or.item(), (loss.item(), loss_or.item())
    num = den = 0.0
    worst = (0.0, None)
    for k, gr in ours.items():
        go = sd_g[k].grad
        if go is None or go.norm().item() < 1e-8:
            assert gr.norm().item() < 1e-5, (k, gr.norm().item())
            continue
        num += ((gr - go) ** 2).sum().item()
        den += (go ** 2).sum().item()
        e = rel_l2(gr, go)
        if e > worst[0]:
            worst = (e, k)
    glob = (num / den) ** 0.5
    print(f"\nsupervised gradient parity: global rel-L2 {glob:.3e} (reference bf16 floor "
          f"{gg['ref_bf16_grad_rel_l2']:.3e}), worst {worst[0]:.3e} {worst[1]}")
    assert glob <= gg["ref_bf16_grad_rel_l2"], glob
    assert worst[0] < 0.25, worst


def test_supervised_trainer_reduces_the_real_image_loss(cuda_device):
    from ifseg_b200.trainer import SegOFATrainer

    g, gg, model, sd, eng, aux, tgt, t2s = _setup()
    _, _, sample = _supervised_sample(g, model, aux, t2s)
    tr = SegOFATrainer(model, lr=2e-4, weight_decay=0.01, clip_norm=1.0, supervised=True, stochastic=False)
    losses = [tr.train_step(sample)["loss"].item() for _ in range(8)]
    assert losses[-1] < losses[0] - 0.05, losses
    assert all(x == x for x in losses)


def test_padded_prompt_gradients_vs_oracle(cuda_device):
    """Prompts of different lengths in one batch (encoder_module.py:730-752): the shorter one is padded; its pad keys are
    masked in the encoder self-attention and the decoder cross-attention, forward and adjoint."""
    from oracle import restated as R

    g, gg, model, sd, eng, aux, tgt, t2s = _setup()
    pad = model.cfg.padding_idx
    aux = dict(aux)
    tok = aux["src_tokens"].clone()
    n = tok.shape[1] // 2
    tok[1, n - 1] = 2  # the second prompt is half as long: eos moves up, pads follow
    tok[1, n:] = pad
    aux["src_tokens"] = tok
    loss, logits = eng.forward_backward(aux, tgt)
    torch.cuda.synchronize()
    ours = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    torch.set_num_threads(8)
    ocfg = oracle_cfg(model.cfg)
    sd_g = {k: (v.clone().requires_grad_() if k in ours else v) for k, v in sd.items()}
    aux_host = {k: v.cpu() for k, v in aux.items()}
    x_or, _ = R.segofa_forward_aux(sd_g, ocfg, aux_host)
    loss_or = R.imfree_loss(x_or, t2s, ocfg)
    loss_or.backward()
    assert abs(loss.item() - loss_or.item()) < 1e-2 * loss_or.item()
    assert rel_l2(logits, x_or.detach()) <= 0.75 * g["ref_bf16_rel_l2"]
    # the mask matters: the same tokens with the pad keys left visible (check_pads=False skips the detection) are off
    logits_nm = eng.forward_train(aux, check_pads=False)["logits"]
    assert rel_l2(logits_nm[1], x_or.detach()[1]) > 3 * rel_l2(logits[1], x_or.detach()[1])
    num = den = 0.0
    for k, gr in ours.items():
        go = sd_g[k].grad
        if go is None or go.norm().item() < 1e-8:
            continue
        num += ((gr - go) ** 2).sum().item()
        den += (go ** 2).sum().item()
    glob = (num / den) ** 0.5
    print(f"\npadded-prompt gradient parity: global rel-L2 {glob:.3e}")
    assert glob <= gg["ref_bf16_grad_rel_l2"], glob
