"""CPU: the criterion's registration boundary (criterions/seg_criterion.py:112-163, 415-597) -- no kernel is launched."""
import types

import pytest
import torch

from helpers import make_task


def test_registered_as_fairseq_criterion_with_dataclass():
    from ifseg_b200 import fairseq_compat as fc
    from ifseg_b200.seg_criterion import SegCriterion, SegCriterionConfig

    assert fc.CRITERION_REGISTRY["seg_criterion"] is SegCriterion and issubclass(SegCriterion, fc.FairseqCriterion)
    cfg = SegCriterionConfig()
    # the schema train.py parses: names, defaults (seg_criterion.py:32-101)
    assert (cfg.label_smoothing, cfg.upscale_lprobs, cfg.unsupervised_segmentation, cfg.criterion_update_freq) == (0.0, "true", "true", 1)
    assert (cfg.init_seg_with_text, cfg.full_context_alignment, cfg.resnet_topk, cfg.resnet_iters) == ("true", "false", 3, 0)
    assert {"report_accuracy", "ignore_prefix_size", "ignore_eos", "sentence_avg", "drop_worst_ratio", "drop_worst_after",
            "use_rdrop", "reg_alpha", "sample_patch_num", "constraint_range", "freeze_embedding_iter",
            "resnet_prob_temperature"} <= set(cfg.__dataclass_fields__)
    cfg.resnet_iters, cfg.label_smoothing, cfg.sentence_avg = 25, 0.1, False
    crit = SegCriterion.build_criterion(cfg, make_task(15))  # fairseq_criterion.py:31-62: arguments looked up by name
    assert isinstance(crit, torch.nn.Module) and crit.resnet_iters == 25 and crit.eps == 0.1
    assert crit.padding_idx == 1 and crit.seg_id_offset == 59457 and crit.iter == -1
    assert SegCriterion.logging_outputs_can_be_summed() is True
    with pytest.raises(ValueError):
        fc.register_criterion("seg_criterion")(SegCriterion)  # duplicate names are refused, as in fairseq's registry


def test_class_name_token_ids_follow_the_reference_encoding():
    from ifseg_b200.seg_criterion import SegCriterion

    task = make_task(3, names=["wall", "traffic light", " sky "])
    ids = SegCriterion(task).class_name_token_ids()
    assert len(ids) == 3 and all(t.dtype == torch.long and t.dim() == 1 for t in ids)
    # ' traffic light' -> bpe(' traffic') + bpe(' light') (seg_criterion.py:375-384)
    enc = lambda w: task.tgt_dict.encode_line(task.bpe.encode(" " + w)).long()  # noqa: E731
    assert torch.equal(ids[1], torch.cat([enc("traffic"), enc("light")])) and torch.equal(ids[2], enc("sky"))


def test_reduce_metrics_aggregates_workers():
    from ifseg_b200 import fairseq_compat as fc
    from ifseg_b200.seg_criterion import SegCriterion

    if not hasattr(fc.metrics, "reset"):
        pytest.skip("real fairseq metrics present")
    fc.metrics.reset()
    C = 4
    g = torch.Generator().manual_seed(0)
    logs = []
    for w in range(3):  # three data-parallel workers
        ai = torch.randint(0, 50, (C,), generator=g).float()
        extra_p, extra_l = torch.randint(0, 20, (C,), generator=g).float(), torch.randint(0, 20, (C,), generator=g).float()
        ap, al = ai + extra_p, ai + extra_l
        d = {"loss": torch.tensor(2.0 + w), "imfree_loss": torch.tensor(2.0 + w), "seg_loss": torch.tensor(1.0), "nll_loss": torch.tensor(1.0),
             "ntokens": 1, "nsentences": 4, "sample_size": 1}
        for sfx in ("", "_lowres", "_resnet_postprocess"):
            d.update({f"area_intersect{sfx}": ai, f"area_pred_label{sfx}": ap, f"area_label{sfx}": al, f"area_union{sfx}": ap + al - ai})
        logs.append(d)
    logs[2]["area_label"][1] = 0.0  # a class absent everywhere in one worker
    SegCriterion.reduce_metrics(logs)
    out = fc.metrics.get_smoothed_values()
    assert out["loss"] == pytest.approx(3.0) and out["nsentences"] == 12 and out["sample_size"] == 3
    assert out["ppl"] == pytest.approx(2.0)
    ai = sum(l["area_intersect"] for l in logs); ap = sum(l["area_pred_label"] for l in logs)
    al = sum(l["area_label"] for l in logs); au = sum(l["area_union"] for l in logs)
    assert out["aAcc"] == round((ai.sum() / ap.sum()).item(), 4)
    assert out["mIoU"] == round(torch.nanmean(ai / au).item(), 4) and out["mAcc"] == round(torch.nanmean(ai / al).item(), 4)
    for sfx in ("_lowres", "_resnet_postprocess"):
        assert {"aAcc" + sfx, "mIoU" + sfx, "mAcc" + sfx} <= set(out)
    assert not any(k.startswith("_") for k in out)


def test_forward_without_cuda_fails_loudly():
    from ifseg_b200.segofa import SegOFAModel
    from ifseg_b200.seg_criterion import SegCriterion

    model = SegOFAModel.from_config("segofa_tiny", 3, 32).eval()
    crit = SegCriterion(make_task(3))
    with pytest.raises(RuntimeError, match="CUDA|no CPU fallback|cuda"):
        crit(model, {"net_input": {}, "target": torch.zeros(1, 2), "ntokens": 1, "nsentences": 1})
