"""-m gpu: END-TO-END parity of the CUDA path at the shapes BASELINE.json quotes (cfg 2, 3, 4, 5), at north_star's
tolerance, through SegOFAModel.forward (C ABI) on the same seeded weights / synthetic inputs as the oracles.

Three gates:
  (T) TRUTH     rel-L2(ours, fp32 oracle restated.py) <= 0.75 x the reference's own bf16-vs-fp32 distance (1.4-1.6e-2);
  (S) SINGLE STORAGE POINT  every kernel launch of a forward at the cfg-2 shape (and a Large one) is replayed on the CPU
                from the launch's own inputs with oracle/matched.py's arithmetic (tests/launch_replay.py): each of the
                ~220 launches must agree to <= 1e-3 rel-L2 -- north_star's "<=1e-3 rel in bf16 logits" at the only
                granularity where it is attainable (measured: 1e-6 ... 3e-4);
  (M) MATCHED, END TO END   rel-L2(ours, oracle/matched.py) <= 1.25 x the oracle's OWN self-distance under a one-fp32-ulp
                jitter of its accumulators.  matched.py is the reference algorithm quantised exactly where the engine
                stores bf16 / fp16 (checked on CPU to be the reference when its quantisers are off).  A quantised chain
                is chaotic (a value crossing a rounding boundary moves by a whole bf16 ulp: d -> sqrt(d * ulp) per storage
                point), so two runs that differ in the last bit of ONE accumulator end 4-8e-3 apart -- that distance,
                not 1e-3, is what any implementation that is not bit-identical in its accumulation order can reach
                end to end (tests/test_oracle_golden.py::test_quantised_chain_is_chaotic pins the fact on CPU).
Masks (argmax of the x16 bilinear upsample of x[:, :-1], seg_criterion.py:237-244, 351):
  * on identical logits the kernel is BIT-EXACT against the oracle (checked here at every shape on OUR logits);
  * end to end against the matched oracle's mask a pixel may differ only where that oracle's top-2 margin is within
    4x the largest logit difference (an argmax cannot be stable below the input difference); the agreement is recorded.
Every measured value is appended to gpurun_out/parity_r02.jsonl (copied to profiles/r02_parity.txt).
"""
import json
import os
import time

import pytest
import torch

from helpers import ROOT, build_cuda_model, load_prompts, oracle_cfg, rel_l2

pytestmark = pytest.mark.gpu

REF_BF16_FLOOR = 1.4e-2  # the reference's own bf16-vs-fp32 rel-L2 (tests/golden/*.pt: 1.35-1.6e-2)
MATCHED_TOL = 1e-3       # north_star: per storage point (launch replay) and for the short image-free / train forwards
JITTER = 1e-7            # ~ one fp32 ulp: the "different accumulation order" model of oracle/matched.py:Q


def record(**kw):
    d = os.path.join(ROOT, "gpurun_out")
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "parity_r02.jsonl"), "a") as f:
        f.write(json.dumps(kw) + "\n")
    print(json.dumps(kw))


def mask_report(model, x, x_matched, hp, wp, S):
    from oracle import restated as R

    eng = model.engine()
    m_dev = eng.predict_mask(x, (hp, wp), (S, S)).view(x.shape[0], -1)
    # identical logits -> bit-identical mask against the reference's own op on this box (F.interpolate on CUDA + argmax)
    xg = x[:, :-1].reshape(x.shape[0], hp, wp, -1).permute(0, 3, 1, 2)
    aten = torch.nn.functional.interpolate(xg, size=(S, S), mode="bilinear", align_corners=False)
    assert torch.equal(m_dev, aten.permute(0, 2, 3, 1).argmax(-1).view(x.shape[0], -1)), "mask differs from ATen CUDA on identical logits"
    m = m_dev.cpu()
    # ... and against the CPU oracle on the same logits up to ATen's own CPU/CUDA rounding difference (near-ties of a few ulps)
    up_same = R.upsample_logits(x.cpu().float(), hp, wp, S, S)[:, :-1]
    t2 = up_same.topk(2, dim=-1).values
    rel_margin = (t2[..., 0] - t2[..., 1]) / t2[..., 0].abs().clamp_min(1e-6)
    assert not ((m != up_same.argmax(-1)) & (rel_margin > 1e-6)).any()
    up = R.upsample_logits(x_matched.float(), hp, wp, S, S)[:, :-1]
    top2 = up.topk(2, dim=-1).values
    margin = top2[..., 0] - top2[..., 1]
    m_or = up.argmax(-1)
    max_err = (x.cpu() - x_matched).abs().max().item()
    dis = m != m_or
    unexplained = int((dis & (margin > 4 * max_err)).sum())
    return dict(same_logits_mask_vs_aten_cuda="bit-exact", same_logits_pixels_differing_from_cpu_oracle=int((m != up_same.argmax(-1)).sum()),
                mask_agree=1.0 - dis.float().mean().item(), mask_pixels=int(dis.numel()), mask_disagree=int(dis.sum()),
                max_abs_logit_err=max_err, disagree_beyond_4x_err=unexplained,
                largest_margin_among_disagreeing=margin[dis].max().item() if dis.any() else 0.0)


CASES = [  # name, arch, classes, image side, batch
    ("cfg2_base480_c15", "segofa_base", 15, 480, 2),
    ("cfg3_base480_c150", "segofa_base", 150, 480, 2),
    ("cfg4_base512_c171", "segofa_base", 171, 512, 1),
    ("cfg5_large640_c150", "segofa_large", 150, 640, 1),
]


@pytest.mark.parametrize("name,arch,C,S,B", CASES)
def test_real_image_branch_at_bench_shapes(cuda_device, name, arch, C, S, B):
    from oracle import matched as M
    from oracle import restated as R
    from ifseg_b200.synthetic import synthetic_inputs

    torch.set_num_threads(max(8, os.cpu_count() or 8))
    model, sd = build_cuda_model(arch, C, S, seed=0)
    inp = synthetic_inputs(model.cfg, B, S, 1, load_prompts()[str(C)])
    oc = oracle_cfg(model.cfg)
    with torch.no_grad():
        t0 = time.time()
        ref, _ = R.segofa_forward(sd, oc, inp["src_tokens"], inp["patch_images"], inp["patch_masks"])
        mt, mextra = M.segofa_forward(sd, oc, inp["src_tokens"], inp["patch_images"], inp["patch_masks"])
        mj, _ = M.segofa_forward(sd, oc, inp["src_tokens"], inp["patch_images"], inp["patch_masks"], jitter=JITTER)
        cpu_s = time.time() - t0
        x, extra = model(**{k: v.cuda() for k, v in inp.items()})
    hp, wp = extra["encoder_returns"]["image_embed_shape"][0]
    assert x.shape == ref.shape == (B, hp * wp + 1, C)
    e_truth, e_matched = rel_l2(x, ref), rel_l2(x, mt)
    e_enc = rel_l2(extra["encoder_returns"]["encoder_out"][0].transpose(0, 1), mextra["encoder_returns"]["encoder_out"])
    e_feat = rel_l2(extra["encoder_returns"]["image_embed_before_proj"][0], mextra["encoder_returns"]["image_embed_before_proj"])
    rep = mask_report(model, x, mt, hp, wp, S)
    record(case=name, branch="real_image", T_e=int(extra["encoder_returns"]["encoder_out"][0].shape[0]), batch=B,
           rel_l2_vs_fp32_oracle=e_truth, rel_l2_vs_matched_oracle=e_matched, stem_vs_matched=e_feat,
           encoder_out_vs_matched=e_enc, oracle_matched_vs_fp32=rel_l2(mt, ref),
           oracle_self_distance_1ulp_jitter=rel_l2(mj, mt), cpu_oracle_seconds=round(cpu_s, 1), **rep)
    assert e_truth <= 0.75 * REF_BF16_FLOOR, e_truth
    assert e_matched <= 1.25 * rel_l2(mj, mt), (e_matched, rel_l2(mj, mt))
    assert rep["disagree_beyond_4x_err"] == 0
    assert rep["mask_agree"] >= 0.98  # 150-171 classes on random-init weights: top-2 margins of 1e-3 are common


@pytest.mark.parametrize("name,arch,C,S,B", [CASES[1], ("cfg1_base128_c15", "segofa_base", 15, 128, 1)])
def test_image_free_branch_at_bench_shapes(cuda_device, name, arch, C, S, B):
    """aux branch (segofa.py:136-151): T_e = 900 + 215 = 1115 at cfg 3; no-grad engine (folded ffn_layernorm) and the
    training engine's forward (separate ffn_layernorm row kernel) against their own matched modes."""
    from oracle import matched as M
    from oracle import restated as R
    from ifseg_b200 import ops

    torch.set_num_threads(max(8, os.cpu_count() or 8))
    model, sd = build_cuda_model(arch, C, S, seed=0)
    prompt = torch.tensor(load_prompts()[str(C)], dtype=torch.long)
    names = torch.full((C, 4), 1, dtype=torch.long)
    lens = torch.zeros(C, dtype=torch.int32)
    g = torch.Generator().manual_seed(4)
    for c in range(C):
        n = 1 + c % 3
        names[c, :n] = torch.randint(4, 50000, (n,), generator=g)
        lens[c] = n
    bag, ends, target = ops.artificial_sample(names.cuda(), lens.cuda(), B, S // 16, S, seed=7)
    aux = dict(src_tokens=prompt.unsqueeze(0).repeat(B, 1), src_lengths=torch.full((B,), prompt.numel()),
               patch_images=bag.cpu(), patch_masks=ends.cpu(), prev_output_tokens=torch.zeros(B, 1, dtype=torch.long))
    oc = oracle_cfg(model.cfg)
    with torch.no_grad():
        ref, _ = R.segofa_forward_aux(sd, oc, aux)
        mt, _ = M.segofa_forward_aux(sd, oc, aux)
        mj, _ = M.segofa_forward_aux(sd, oc, aux, jitter=JITTER)
        mt_train, _ = M.segofa_forward_aux(sd, oc, aux, fold_ffn=False)
        _, extra = model(aux_input={k: v.cuda() for k, v in aux.items()})
        x = extra["aux_output"][0]
    model.train()
    te = model.train_engine()
    te.stochastic = False
    xt = te.forward_train({k: v.cuda() for k, v in aux.items()})["logits"]
    model.eval()
    e_truth, e_matched, e_train = rel_l2(x, ref), rel_l2(x, mt), rel_l2(xt, mt_train)
    record(case=name, branch="image_free", T_e=(S // 16) ** 2 + prompt.numel(), batch=B, rel_l2_vs_fp32_oracle=e_truth,
           rel_l2_vs_matched_oracle=e_matched, train_forward_vs_matched_oracle=e_train,
           train_forward_vs_fp32_oracle=rel_l2(xt, ref), oracle_self_distance_1ulp_jitter=rel_l2(mj, mt))
    floor = rel_l2(mj, mt)
    assert e_truth <= 0.75 * REF_BF16_FLOOR
    assert e_matched <= 1.25 * floor, (e_matched, floor)
    assert e_train <= 1.25 * floor, (e_train, floor)


def test_train_step_gradients_at_the_cfg3_shape(cuda_device):
    """Gradient parity where the train step is quoted (cfg 3: Base, 480x480, 150 classes, T_e = 1115), batch 2 to bound
    the CPU oracle's autograd: image-free forward + compute_imfree_loss + backward of the training engine against the
    fp32 oracle.  Tolerance: the reference's own bf16-vs-fp32 gradient distance (4.8e-2 global rel-L2, stored in
    tests/golden/golden_grads_base_c150_s64_b2.pt)."""
    from oracle import restated as R
    from ifseg_b200 import ops
    from ifseg_b200.seg_criterion import class_targets

    name, arch, C, S, B = CASES[1]
    torch.set_num_threads(max(8, os.cpu_count() or 8))
    model, sd = build_cuda_model(arch, C, S, seed=0)
    prompt = torch.tensor(load_prompts()[str(C)], dtype=torch.long)
    names = torch.full((C, 4), 1, dtype=torch.long)
    lens = torch.zeros(C, dtype=torch.int32)
    g = torch.Generator().manual_seed(4)
    for c in range(C):
        n = 1 + c % 3
        names[c, :n] = torch.randint(4, 50000, (n,), generator=g)
        lens[c] = n
    bag, ends, target = ops.artificial_sample(names.cuda(), lens.cuda(), B, S // 16, S, seed=7)
    aux = dict(src_tokens=prompt.unsqueeze(0).repeat(B, 1), src_lengths=torch.full((B,), prompt.numel()),
               patch_images=bag.cpu(), patch_masks=ends.cpu(), prev_output_tokens=torch.zeros(B, 1, dtype=torch.long))
    t2s = target.cpu().long()
    model.train()
    te = model.train_engine()
    te.stochastic = False
    tgt = class_targets(t2s[:, :-1].reshape(B, S, S), 59457, C).cuda()
    loss, _ = te.forward_backward({k: v.cuda() for k, v in aux.items()}, tgt)
    torch.cuda.synchronize()
    ours = {k: p.grad.detach().float().cpu() for k, p in model.named_parameters() if p.grad is not None}
    t0 = time.time()
    oc = oracle_cfg(model.cfg)
    sd_g = {k: (v.clone().requires_grad_() if k in ours else v) for k, v in sd.items()}
    x_or, _ = R.segofa_forward_aux(sd_g, oc, aux)
    loss_or = R.imfree_loss(x_or, t2s, oc)
    loss_or.backward()
    cpu_s = time.time() - t0
    num = den = 0.0
    per, worst = [], (0.0, None)
    for k, gr in ours.items():
        go = sd_g[k].grad
        if go is None or go.norm().item() < 1e-8:
            assert gr.norm().item() < 1e-5, (k, gr.norm().item())
            continue
        num += ((gr - go) ** 2).sum().item()
        den += (go ** 2).sum().item()
        e = rel_l2(gr, go)
        per.append(e)
        if e > worst[0]:
            worst = (e, k)
    per.sort()
    glob = (num / den) ** 0.5
    record(case=name, branch="image_free_gradients", T_e=(S // 16) ** 2 + prompt.numel(), batch=B, tensors=len(per),
           loss=loss.item(), oracle_loss=loss_or.item(), grad_global_rel_l2_vs_fp32_oracle=glob,
           grad_per_tensor_median=per[len(per) // 2], grad_per_tensor_worst=worst[0], worst_tensor=worst[1],
           reference_bf16_grad_floor=4.8e-2, cpu_oracle_seconds=round(cpu_s, 1))
    assert abs(loss.item() - loss_or.item()) < 1e-2 * loss_or.item()
    assert len(ours) == 380
    assert glob <= 4.8e-2, glob
    assert worst[0] < 0.25, worst


@pytest.mark.parametrize("name,arch,C,S", [("cfg2_base480_c15", "segofa_base", 15, 480), ("large320_c150", "segofa_large", 150, 320)])
def test_every_launch_replayed_from_its_own_inputs(cuda_device, name, arch, C, S):
    """Gate (S): the whole forward (stem, position bias, encoder, decoder, head) at batch 1, every GEMM / convolution /
    attention / row-kernel launch compared with the matched oracle's arithmetic on the SAME inputs."""
    from launch_replay import ReplayTap
    from ifseg_b200 import ops
    from ifseg_b200.synthetic import synthetic_inputs

    torch.set_num_threads(max(8, os.cpu_count() or 8))
    model, sd = build_cuda_model(arch, C, S, seed=0)
    inp = synthetic_inputs(model.cfg, 1, S, 1, load_prompts()[str(C)])
    tap = ReplayTap()
    ops.set_tap(tap)
    try:
        with torch.no_grad():
            model(**{k: v.cuda() for k, v in inp.items()})
    finally:
        ops.set_tap(None)
    summ = tap.summary()
    record(case=name, branch="launch_replay", launches=len(tap.records), **{k: v for k, v in summ.items()})
    assert len(tap.records) > 150
    for r in tap.records:
        assert r["rel_l2"] <= MATCHED_TOL, r
        assert r.get("rowstats_rel_l2", 0.0) <= 1e-4, r
