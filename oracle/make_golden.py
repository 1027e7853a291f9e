"""TEST INFRASTRUCTURE -- run in the BUILD container only (needs /root/reference).

1. PINS oracle/restated.py: builds the UNMODIFIED reference SegOFAModel (oracle/ref_shim.py),
   loads the deterministic synthetic state dict (ifseg_b200.synthetic.generate_state_dict), runs
   the reference forward on seeded inputs and asserts the restatement reproduces it to fp32
   round-off, for the real-image branch, the image-free (aux) branch, the non-causal branch
   and a padded batch.
2. Emits the golden fixtures (outputs OF THE REFERENCE ITSELF) under tests/golden/:
     manifest_<cfg>.json   -- names/shapes/dtypes of the reference state dict (checkpoint compat)
     golden_<cfg>.pt       -- inputs, reference logits (fp32), a few intermediates, reference-bf16
                              noise floor; small enough to commit
     prompts.json          -- BPE ids of the real prompts (15/150/171 classes) from utils/BPE
   The GPU box has no /root/reference: tests there regenerate the weights from the same seeded
   generator and compare against these files.

Usage:  python oracle/make_golden.py            (about 1 minute on 8 cores)
"""
import json
import os
import re
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from ifseg_b200.config import preset  # noqa: E402
from ifseg_b200.synthetic import generate_state_dict, synthetic_inputs  # noqa: E402
from oracle import restated as R  # noqa: E402
from oracle.ref_shim import REF_ROOT, build_reference_model  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")


def oracle_cfg(cfg):
    return R.SegOFAConfig(**{k: getattr(cfg, k) for k in R.SegOFAConfig.__dataclass_fields__ if hasattr(cfg, k)})


def build_prompts():
    """Real prompt token ids (segmentation_dataset.py:175-186, 272-281) with the reference BPE."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(
        "gpt2_bpe_utils", os.path.join(REF_ROOT, "custom_fairseq/fairseq/data/encoders/gpt2_bpe_utils.py"))
    bpe_utils = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bpe_utils)
    enc = bpe_utils.get_encoder(os.path.join(REF_ROOT, "utils/BPE/encoder.json"), os.path.join(REF_ROOT, "utils/BPE/vocab.bpe"))
    sym2idx = {}
    with open(os.path.join(REF_ROOT, "utils/BPE/dict.txt")) as f:
        for i, line in enumerate(f):
            sym2idx[line.rsplit(" ", 1)[0]] = i + 4  # <s>, <pad>, </s>, <unk> first

    def encode_text(text):  # segmentation_dataset.py:193-200
        words = [" ".join(map(str, enc.encode(" {}".format(wd.strip())))) for wd in text.strip().split()]
        return [sym2idx.get(t, 3) for t in " ".join(words).split()]

    out = {}
    for script in ("coco_unseen.sh", "ade.sh", "coco_fine.sh"):
        s = open(os.path.join(REF_ROOT, "run_scripts/IFSeg", script)).read()
        cats = re.search(r"^category_list='(.*)'$", s, re.M).group(1)
        prefix = re.search(r"^prompt_prefix='(.*)'$", s, re.M).group(1)
        names = [x.strip() for x in cats.split(",")] + ["unknown"]
        ids = [0] + encode_text(f" {prefix.lstrip()}")
        for n in names:
            ids += encode_text(f" {n}")
        ids += [2]
        out[str(len(names) - 1)] = ids
    return out


def rel(a, b):
    return ((a - b).norm() / b.norm()).item()


def run_case(name, arch, num_seg, size, batch, prompts, save=True):
    cfg = preset(arch, num_seg=num_seg, patch_image_size=size, orig_patch_image_size=size)
    ref, _ = build_reference_model(arch, num_seg, size)
    sd = generate_state_dict(cfg, seed=0)
    missing = ref.load_state_dict(sd, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    ref.eval()
    manifest = [[k, list(v.shape), str(v.dtype).replace("torch.", "")] for k, v in ref.state_dict().items()]
    inp = synthetic_inputs(cfg, batch, size, seed=1, src_tokens=prompts[str(num_seg)])
    ocfg = oracle_cfg(cfg)
    with torch.no_grad():
        x_ref, extra_ref = ref(**inp)
        x_or, extra_or = R.segofa_forward(sd, ocfg, inp["src_tokens"], inp["patch_images"], inp["patch_masks"],
                                          inp["prev_output_tokens"], want_attn=True)
        err = (x_ref - x_or).abs().max().item()
        print(f"[{name}] real-image: logits {tuple(x_ref.shape)} max|ref-oracle| = {err:.3e}  (|logit|max {x_ref.abs().max():.3f})")
        assert err < 5e-5 * max(1.0, x_ref.abs().max().item()), "restated oracle diverges from the reference"
        enc_ref = extra_ref["encoder_returns"]["encoder_out"][0].transpose(0, 1)
        assert (enc_ref - extra_or["encoder_returns"]["encoder_out"]).abs().max() < 5e-5
        assert (extra_ref["attn"][0] - extra_or["attn"]).abs().max() < 1e-5
        feat_ref = extra_ref["encoder_returns"]["image_embed_before_proj"][0]
        assert (feat_ref - extra_or["encoder_returns"]["image_embed_before_proj"]).abs().max() < 1e-4 * feat_ref.abs().max()
        # non-causal decoder
        x_ref_fc, _ = ref(**inp, full_context_alignment=True)
        x_or_fc, _ = R.segofa_forward(sd, ocfg, inp["src_tokens"], inp["patch_images"], inp["patch_masks"],
                                      full_context_alignment=True)
        assert (x_ref_fc - x_or_fc).abs().max() < 5e-5 * max(1.0, x_ref_fc.abs().max().item())
        # padded text (has_pads branch): pad the tail of sample 0's prompt
        inp_p = {k: v.clone() for k, v in inp.items()}
        inp_p["src_tokens"][0, -5:] = 1
        x_ref_p, _ = ref(**inp_p)
        x_or_p, _ = R.segofa_forward(sd, ocfg, inp_p["src_tokens"], inp_p["patch_images"], inp_p["patch_masks"])
        assert (x_ref_p - x_or_p).abs().max() < 5e-5 * max(1.0, x_ref_p.abs().max().item())
        # image-free branch (aux_input), rand_k-1-33 style grid (segmentation_dataset.py:303-329)
        g = torch.Generator().manual_seed(2)
        hp = size // 16
        labels = torch.randint(0, num_seg, (batch, hp * hp), generator=g)
        name_ids = []
        for n in range(num_seg):  # tokens of class n: reuse slices of the prompt as stand-in bags of 1-3 tokens
            ln = 1 + n % 3
            name_ids.append(inp["src_tokens"][0, 13 + n: 13 + n + ln])
        bags, offs = [], []
        for b in range(batch):
            toks = [name_ids[int(l)] for l in labels[b]]
            lens = torch.tensor([len(t) for t in toks])
            bags.append(torch.cat(toks))
            offs.append(lens.cumsum(0))
        L = max(len(t) for t in bags)
        bag_tokens = torch.full((batch, L), 1, dtype=torch.long)
        for b in range(batch):
            bag_tokens[b, : len(bags[b])] = bags[b]
        aux = dict(src_tokens=inp["src_tokens"], src_lengths=inp["src_lengths"], patch_images=bag_tokens,
                   patch_masks=torch.cat(offs), prev_output_tokens=inp["prev_output_tokens"])
        _, extra_aux = ref(aux_input=aux)
        x_aux_ref = extra_aux["aux_output"][0]
        x_aux_or, _ = R.segofa_forward_aux(sd, ocfg, aux)
        e2 = (x_aux_ref - x_aux_or).abs().max().item()
        print(f"[{name}] image-free : max|ref-oracle| = {e2:.3e}")
        assert e2 < 5e-5 * max(1.0, x_aux_ref.abs().max().item())
        # reference noise floor: the reference's own bf16 run vs its fp32 run
        ref_bf16 = ref.to(torch.bfloat16)
        inp16 = dict(inp)
        inp16["patch_images"] = inp["patch_images"].to(torch.bfloat16)
        x_bf16, _ = ref_bf16(**inp16)
        floor = rel(x_bf16.float(), x_ref)
        x_emul, _ = R.segofa_forward(sd, ocfg, inp["src_tokens"], inp["patch_images"], inp["patch_masks"], bf16=True)
        print(f"[{name}] reference bf16-vs-fp32 rel-L2 = {floor:.3e}; restated(emulate_bf16) rel-L2 = {rel(x_emul, x_ref):.3e}")
        mask_ref = R.predict_mask(x_ref, hp, hp, size, size)
        top2 = x_ref[:, :-1].topk(2, dim=-1).values
        margin = (top2[..., 0] - top2[..., 1])
        print(f"[{name}] min top-2 margin {margin.min():.3e}, argmax agreement ref-bf16 vs fp32 (low-res): "
              f"{(x_bf16[:, :-1].float().argmax(-1) == x_ref[:, :-1].argmax(-1)).float().mean():.4f}")
    if save:
        with open(os.path.join(GOLD, f"manifest_{name}.json"), "w") as f:
            json.dump(manifest, f)
        torch.save(dict(
            arch=arch, num_seg=num_seg, image_size=size, batch=batch, weight_seed=0, input_seed=1,
            src_tokens=inp["src_tokens"], logits=x_ref.clone(), logits_full_context=x_ref_fc.clone(),
            logits_padded=x_ref_p.clone(), encoder_out=enc_ref.clone(), resnet_features=feat_ref.clone(),
            attn=extra_ref["attn"][0].clone(), mask=mask_ref.to(torch.int16), ref_bf16_rel_l2=floor,
            aux_input={k: v for k, v in aux.items()}, aux_logits=x_aux_ref.clone(),
        ), os.path.join(GOLD, f"golden_{name}.pt"))
    return floor


def main():
    os.makedirs(GOLD, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    prompts = build_prompts()
    print({k: len(v) for k, v in prompts.items()})
    assert {k: len(v) for k, v in prompts.items()} == {"15": 36, "150": 215, "171": 239}, "T_txt differs from SURVEY s8"
    with open(os.path.join(GOLD, "prompts.json"), "w") as f:
        json.dump(prompts, f)
    # cfg 1 of BASELINE.json (Base, 128x128, 15 classes, B=1) + a 2-image batch at 150 classes
    run_case("base_c15_s128", "segofa_base", 15, 128, 1, prompts)
    run_case("base_c150_s64_b2", "segofa_base", 150, 64, 2, prompts)
    # the other model families (the restated oracle is pinned on the reference for each of them)
    run_case("tiny_c15_s64", "segofa_tiny", 15, 64, 1, prompts)
    run_case("medium_c15_s64", "segofa_medium", 15, 64, 1, prompts)
    run_case("large_c15_s64", "segofa_large", 15, 64, 1, prompts)
    run_case("huge_c15_s64", "segofa_huge", 15, 64, 1, prompts)
    run_grad_case("base_c150_s64_b2")
    run_grad_case("base_c15_s128")  # cfg 1 shape (B=1, 15 classes): second pin of the oracle's autograd
    if "--full" in sys.argv:  # cfg 2 shape (not saved: 8x901x15 logits only) -- minutes of CPU
        run_case("base_c15_s480", "segofa_base", 15, 480, 2, prompts, save=False)


if __name__ == "__main__":
    main()


def run_grad_case(name="base_c150_s64_b2"):
    """Gradient fixture of the image-free training branch, generated by the UNMODIFIED reference model
    (eval mode: dropout / DropPath off -- the deterministic gradient-parity configuration of SURVEY.md s8d)
    + compute_imfree_loss (seg_criterion.py:246-267 with 32/512 generalised), autograd backward.
    Saves the per-parameter gradient norms of every tensor that receives a gradient, a few small gradients
    in full, the loss and the target; asserts that the restated oracle's autograd reproduces all of it."""
    g = torch.load(os.path.join(GOLD, f"golden_{name}.pt"))
    arch, num_seg, size, batch = g["arch"], g["num_seg"], g["image_size"], g["batch"]
    cfg = preset(arch, num_seg=num_seg, patch_image_size=size, orig_patch_image_size=size)
    ocfg = oracle_cfg(cfg)
    ref, _ = build_reference_model(arch, num_seg, size)
    sd = generate_state_dict(cfg, seed=0)
    ref.load_state_dict(sd, strict=True)
    ref.eval()
    gen = torch.Generator().manual_seed(5)
    # dictionary-id targets on the full-resolution grid (+ trailing eos), incl. the ignored "unknown" id
    t2s = torch.cat([torch.randint(0, num_seg + 1, (batch, size * size), generator=gen) + 59457,
                     torch.full((batch, 1), 2)], 1)
    aux = g["aux_input"]
    _, extra = ref(aux_input=aux)
    loss_ref = R.imfree_loss(extra["aux_output"][0], t2s, ocfg)
    ref.zero_grad()
    loss_ref.backward()
    grads_ref = {k: p.grad.clone() for k, p in ref.named_parameters() if p.grad is not None}
    # the restated oracle under autograd
    sd_g = {k: (v.clone().requires_grad_() if v.is_floating_point() and k in grads_ref else v) for k, v in sd.items()}
    x_or, _ = R.segofa_forward_aux(sd_g, ocfg, aux)
    loss_or = R.imfree_loss(x_or, t2s, ocfg)
    loss_or.backward()
    worst = 0.0
    for k, gr in grads_ref.items():
        go = sd_g[k].grad
        if go is None or gr.norm().item() < 1e-8:  # k_proj / pos_k biases: exactly zero in exact arithmetic
            continue                                 # (softmax is invariant to a per-row constant): fp noise only
        worst = max(worst, ((go - gr).norm() / gr.norm().clamp_min(1e-30)).item())
    print(f"[grads {name}] loss ref {loss_ref.item():.6f} oracle {loss_or.item():.6f}; {len(grads_ref)} tensors with "
          f"gradients; worst rel-L2(oracle autograd, reference autograd) = {worst:.3e}")
    assert abs(loss_ref.item() - loss_or.item()) < 1e-5 and worst < 2e-4
    # noise floor: the reference's own bf16 run (model.to(bf16), trainer.py:95-101) against its fp32 run
    ref16 = ref.to(torch.bfloat16)
    ref16.zero_grad()
    _, extra16 = ref16(aux_input=aux)
    loss16 = R.imfree_loss(extra16["aux_output"][0], t2s, ocfg)
    loss16.backward()
    g16 = {k: p.grad.float() for k, p in ref16.named_parameters() if p.grad is not None}
    floor = {k: ((g16[k] - v).norm() / v.norm().clamp_min(1e-30)).item() for k, v in grads_ref.items() if k in g16}
    num = sum(((g16[k] - v) ** 2).sum().item() for k, v in grads_ref.items() if k in g16)
    den = sum((v ** 2).sum().item() for k, v in grads_ref.items() if k in g16)
    floor_global = (num / den) ** 0.5
    fl = sorted(v for k, v in floor.items() if grads_ref[k].norm().item() > 1e-8)
    print(f"[grads {name}] reference bf16-vs-fp32: loss {loss16.item():.4f}; global grad rel-L2 {floor_global:.3e}; "
          f"per-tensor median {fl[len(fl) // 2]:.3e} max {fl[-1]:.3e}")
    small = {k: v for k, v in grads_ref.items() if v.numel() <= 4096 and ("layers.0." in k or "layers.5." in k or ".layers." not in k)}
    torch.save(dict(name=name, loss=loss_ref.item(), text2seg_target=t2s.to(torch.int32),
                    grad_norms={k: v.norm().item() for k, v in grads_ref.items()},
                    grad_numel={k: v.numel() for k, v in grads_ref.items()}, small_grads=small,
                    ref_bf16_grad_rel_l2=floor_global, ref_bf16_grad_rel_l2_per_tensor=floor, ref_bf16_loss=loss16.item()),
               os.path.join(GOLD, f"golden_grads_{name}.pt"))
    return grads_ref
