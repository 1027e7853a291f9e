"""`seg_criterion`, B200-native.  Mirrors criterions/seg_criterion.py of alinlab/ifseg:

  registration                                           -> @register_criterion("seg_criterion", dataclass=SegCriterionConfig),
                                                            a FairseqCriterion with the reference's constructor arguments,
                                                            forward contract, reduce_metrics and logging_outputs_can_be_summed
                                                            (:112-163, 415-597), so train.py / trainer.py / tasks build and
                                                            drive it unchanged;
  upsample_logits + compute_metric (:237-244, 349-362)  -> one fused kernel (sgf_upsample_argmax): the
      [B,C,H,W] logits the reference materialises (138 MB / image at C=150) never exist;
  compute_imfree_loss / the display CE (:246-267, 340)  -> sgf_upsample_ce_loss (+ sgf_upsample_ce_loss_bwd);
  _lazy_initialization (:373-407)                       -> sgf_embedding_bag_mean, run automatically on the first forward;
  SegCriterion.forward (:165-235)                       -> same signature and logging_output keys, incl. the `_lowres`
                                                            and `_resnet_postprocess` metric families.

The training branch (unsupervised_segmentation: image-free loss + no-grad real-image metrics) returns a loss whose
.backward() runs the hand-written adjoint kernels (PixelCrossEntropyFunction -> train_engine.ImFreeBranchFunction).
"""
import math

import torch

from dataclasses import dataclass, field
from typing import Optional

from . import ops
from .fairseq_compat import FairseqCriterion, FairseqDataclass, metrics, register_criterion, utils
from .segofa import str_bool

try:  # pragma: no cover - real fairseq deployment: sentence_avg is interpolated from the optimization config
    from omegaconf import II

    _SENTENCE_AVG = II("optimization.sentence_avg")
except Exception:
    _SENTENCE_AVG = False


def class_targets(target_ids, seg_id_offset, num_seg, padding_idx=1):
    """dictionary ids -> class ids; pad and the 'unknown' class (id seg_id_offset+num_seg) become -1
    (ignored), as the masks at seg_criterion.py:259, 306-311 do."""
    t = target_ids - seg_id_offset
    bad = (target_ids == padding_idx) | (target_ids == seg_id_offset + num_seg) | (t < 0) | (t >= num_seg)
    return torch.where(bad, torch.full_like(t, -1), t)


def segmentation_metrics(logits, target_classes, hp, wp):
    """(mask [B,h,w] int64, area_intersect, area_pred_label, area_label, area_union) -- compute_metric."""
    mask, areas = ops.upsample_argmax(logits.float().contiguous(), hp, wp, target_classes.shape[1],
                                      target_classes.shape[2], target=target_classes.contiguous())
    return mask, areas[0], areas[1], areas[2], areas[1] + areas[2] - areas[0]


def pixel_cross_entropy(logits, target_classes, hp, wp, label_smoothing=0.0):
    """F.cross_entropy(upsample(logits), target) over the non-ignored pixels, fused."""
    loss, _ = ops.upsample_ce_loss(logits.float().contiguous(), target_classes.contiguous(), hp, wp, label_smoothing)
    return loss


class PixelCrossEntropyFunction(torch.autograd.Function):
    """compute_imfree_loss (seg_criterion.py:246-267) as one autograd node: forward = sgf_upsample_ce_loss (the
    up-sampled logits never exist), backward = sgf_upsample_ce_loss_bwd (gradient w.r.t. the patch-grid logits)."""

    @staticmethod
    def forward(ctx, logits, target_classes, hp, wp, eps):
        logits = logits.float().contiguous()
        tgt = target_classes.contiguous()
        lse = torch.empty(tuple(tgt.shape), dtype=torch.float32, device=logits.device)
        acc = ops.upsample_ce_loss(logits, tgt, hp, wp, eps, lse_out=lse, raw=True)
        ctx.save_for_backward(logits, tgt, lse, acc)
        ctx.hw, ctx.eps = (hp, wp), eps
        return acc[0] / acc[1]

    @staticmethod
    def backward(ctx, gout):
        logits, tgt, lse, acc = ctx.saved_tensors
        B, T, C = logits.shape
        dl = torch.empty((B, T, (C + 7) // 8 * 8), dtype=torch.bfloat16, device=logits.device)
        ops.upsample_ce_loss_bwd(logits, tgt, lse, acc[1:], ctx.hw[0], ctx.hw[1], dl, ctx.eps, 1.0)
        return dl[..., :C].float() * gout, None, None, None, None


def derive_metrics(area_intersect, area_pred_label, area_label, area_union):
    """aAcc / mIoU / mAcc exactly as reduce_metrics derives them (seg_criterion.py:533-572)."""
    aacc = (area_intersect.sum() / area_pred_label.sum()).item()
    miou = torch.nanmean(area_intersect / area_union).item()
    macc = torch.nanmean(area_intersect / area_label).item()
    return dict(aAcc=round(aacc, 4), mIoU=round(miou, 4), mAcc=round(macc, 4))


def _f(default, help):  # noqa: A002
    return field(default=default, metadata={"help": help})


@dataclass
class SegCriterionConfig(FairseqDataclass):
    """Command-line schema of the criterion (criterions/seg_criterion.py:32-101): same names, types and defaults."""

    label_smoothing: float = _f(0.0, "epsilon for label smoothing, 0 means no label smoothing")
    report_accuracy: bool = _f(False, "report accuracy metric")
    ignore_prefix_size: int = _f(0, "Ignore first N tokens")
    ignore_eos: bool = _f(True, "Ignore eos token")
    sentence_avg: bool = _SENTENCE_AVG
    drop_worst_ratio: float = _f(0.0, "ratio for discarding bad samples")
    drop_worst_after: int = _f(0, "steps for discarding bad samples")
    use_rdrop: bool = _f(False, "use R-Drop")
    reg_alpha: float = _f(1.0, "weight for R-Drop")
    sample_patch_num: int = _f(196, "sample patches for v1")
    constraint_range: Optional[str] = _f(None, "constraint range")
    upscale_lprobs: str = _f("true", "true | false")
    unsupervised_segmentation: str = _f("true", "true | false")
    criterion_update_freq: int = _f(1, "update frequency used in this criterion")
    freeze_embedding_iter: int = _f(-1, "freeze the token embedding after this (effective) iteration; -1 = never")
    full_context_alignment: str = _f("false", "whether to apply full attention in decoder")
    init_seg_with_text: str = _f("true", "whether to lazy initialize the segmentation with text embedding bags")
    resnet_topk: int = _f(3, "filtering with topk adjacent resnet features")
    resnet_prob_temperature: float = _f(1.0, "resnet softmax temperature")
    resnet_iters: int = _f(0, "resnet filtering iterations")


_METRIC_FAMILIES = ("", "_lowres", "_resnet_postprocess")  # suffixes of the area_* logging keys (:323-337)


@register_criterion("seg_criterion", dataclass=SegCriterionConfig)
class SegCriterion(FairseqCriterion):
    """Same constructor arguments / forward contract as the reference criterion (seg_criterion.py:115-163).  `task` needs
    `target_dictionary` (`pad()`, `eos()`, `index("<seg_0>")`), `.cfg.num_seg_tokens`, `.cfg.category_list`, and -- for
    init_seg_with_text -- `.bpe.encode` and `.tgt_dict.encode_line` (tasks/ofa_task.py)."""

    def __init__(self, task, sentence_avg=False, label_smoothing=0.0, ignore_prefix_size=0, ignore_eos=True,
                 report_accuracy=False, drop_worst_ratio=0, drop_worst_after=0, use_rdrop=False, reg_alpha=1.0,
                 sample_patch_num=196, constraint_range=None, upscale_lprobs="true", unsupervised_segmentation="true",
                 criterion_update_freq=1, freeze_embedding_iter=-1, full_context_alignment="false",
                 init_seg_with_text="true", resnet_topk=3, resnet_prob_temperature=1.0, resnet_iters=0):
        super().__init__(task)
        self.sentence_avg = sentence_avg
        self.eps = label_smoothing
        self.sample_patch_num = sample_patch_num
        self.upscale_lprobs = str_bool(upscale_lprobs)
        self.unsupervised_segmentation = str_bool(unsupervised_segmentation)
        self.full_context_alignment = str_bool(full_context_alignment)
        self.init_seg_with_text = str_bool(init_seg_with_text)
        self.resnet_topk, self.resnet_prob_temperature, self.resnet_iters = resnet_topk, resnet_prob_temperature, resnet_iters
        self.criterion_update_freq = criterion_update_freq
        self.iter = -1            # -1 = "first forward still to come": restores the counter and runs the lazy init
        self.effective_iter = -1  # iter // criterion_update_freq
        if not hasattr(self, "padding_idx"):
            self.padding_idx = task.target_dictionary.pad()
        self.seg_id_offset = task.target_dictionary.index("<seg_0>")
        self.eos_idx = task.target_dictionary.eos() if hasattr(task.target_dictionary, "eos") else 2
        self.num_seg = task.cfg.num_seg_tokens
        self.id2rawtext = [x.strip() for x in task.cfg.category_list.split(",")]
        assert len(self.id2rawtext) == self.num_seg

    # -- seg_criterion.py:373-407 -----------------------------------------------------------
    def class_name_token_ids(self):
        """BPE ids of ' <class name>' for every class, through the task's own encoder (seg_criterion.py:375-384)."""
        bpe, d = self.task.bpe, self.task.tgt_dict

        def encode(text):
            line = " ".join(bpe.encode(" {}".format(w.strip())) for w in text.strip().split())
            return d.encode_line(line=line, add_if_not_exist=False, append_eos=False).long()

        return [encode(f" {name}") for name in self.id2rawtext]

    def _lazy_initialization(self, sample, model, ema_model=None):
        if self.init_seg_with_text:
            self.lazy_initialization(model, self.class_name_token_ids(), ema_model=ema_model)

    def lazy_initialization(self, model, class_token_ids, ema_model=None):
        """seg_embed_tokens (and seg_projection) <- mean token embedding of each class name, for `model` and `ema_model`.
        class_token_ids: list of 1-D LongTensors, one per class.  The parameters are written IN PLACE, so the ties
        (encoder/decoder seg_embed_tokens, tied seg_projection), the training engine's arena views and external
        optimizers keep pointing at the same storage; every derived device copy is invalidated."""
        dev = model.encoder.embed_tokens.weight.device
        lens = torch.tensor([len(t) for t in class_token_ids], dtype=torch.long)
        tokens = torch.cat(class_token_ids).to(dev).unsqueeze(0)
        table = model.encoder.embed_tokens.weight.detach()
        if table.dtype not in (torch.float32, torch.bfloat16):
            table = table.float()
        avg = ops.embedding_bag_mean(tokens, lens.cumsum(0).to(dev), table.contiguous(), len(lens))
        for m in (model, ema_model):
            if m is None:
                continue
            with torch.no_grad():
                m.encoder.seg_embed_tokens.weight.data.copy_(avg)
                m.decoder.seg_embed_tokens.weight.data.copy_(avg)
                if not m.decoder.tie_seg_projection:
                    m.decoder.seg_projection.weight.data.copy_(avg)
            if hasattr(m, "parameters_changed"):
                m.parameters_changed()

    def _update_iteration(self):
        self.iter += 1
        self.effective_iter = self.iter // self.criterion_update_freq

    # -- seg_criterion.py:165-235 -------------------------------------------------------------
    def forward(self, model, sample, update_num=0, reduce=True, ema_model=None):
        if self.iter == -1:  # first call (also after a restart: update_num restores the counter), :173-176
            self.iter = self.criterion_update_freq * update_num - 1
            self._lazy_initialization(sample, model, ema_model)
        self._update_iteration()
        ntokens = 1  # compute_loss returns ntokens = 1 (:345)
        if model.training and self.unsupervised_segmentation:  # seg_criterion.py:178-186
            S = model.cfg.patch_image_size
            net_output = model(full_context_alignment=self.full_context_alignment, aux_input=sample["aux_input"])
            logits = net_output[1]["aux_output"][0]
            ids = sample["text2seg_target"][:, :-1].reshape(-1, S, S).to(logits.device)
            tgt = class_targets(ids, self.seg_id_offset, self.num_seg, self.padding_idx)
            imfree_loss = PixelCrossEntropyFunction.apply(logits, tgt, S // 16, S // 16, self.eps)
            loss = imfree_loss
            with torch.no_grad():
                seg_logits, seg_extra = model(**sample["net_input"], full_context_alignment=self.full_context_alignment)
                seg_loss, metric_out = self.compute_loss(seg_logits, seg_extra, sample, training=True)
        elif model.training:  # supervised: the real-image loss carries the gradient (seg_criterion.py:188-192)
            logits, extra = model(**sample["net_input"], full_context_alignment=self.full_context_alignment)
            seg_loss, metric_out = self.compute_loss(logits, extra, sample, training=True)
            imfree_loss = torch.zeros(1, device=logits.device)
            loss = seg_loss
        else:
            with torch.no_grad():
                logits, extra = model(**sample["net_input"], full_context_alignment=self.full_context_alignment)
                if self.resnet_iters > 0:  # seg_criterion.py:197-213: label propagation over ResNet-feature neighbours
                    feats = extra["encoder_returns"]["image_embed_before_proj"][0]
                    prob, _ = ops.label_propagation(feats.contiguous(), logits.float().contiguous(), self.resnet_topk,
                                                    self.resnet_iters, self.resnet_prob_temperature)
                    eos = prob.new_zeros((prob.shape[0], 1, prob.shape[2]))  # fake eos row (:211)
                    extra["resnet_postprocess_probability"] = torch.cat([prob, eos], dim=1)
                seg_loss, metric_out = self.compute_loss(logits, extra, sample, training=model.training)
            imfree_loss = torch.zeros(1, device=logits.device)
            loss = seg_loss
        sample_size = sample["target"].size(0) if self.sentence_avg else ntokens
        logging_output = {"loss": loss.data, "imfree_loss": imfree_loss.data, "seg_loss": seg_loss.data,
                          "ntokens": sample["ntokens"], "nsentences": sample["nsentences"], "sample_size": sample_size}
        for k, v in metric_out.items():
            logging_output[k] = v.data if torch.is_tensor(v) else v
        return loss, sample_size, logging_output

    # -- seg_criterion.py:269-347 ---------------------------------------------------------------
    def compute_loss(self, logits, extra, sample, training=False):
        """Metrics (and the display CE) of one real-image forward.  Training: targets at the network input resolution
        (`sample["target"]`); evaluation: the original-resolution ground truth (`ori_semantic_seg`, batch 1) when the
        sample carries it (:283-287).  `downsampled_target` adds the `_lowres` family: argmax of the patch-grid logits
        against patch-grid labels (:273-281, 314-321)."""
        hp, wp = extra["encoder_returns"]["image_embed_shape"][0]
        dev = logits.device
        if not training and sample.get("ori_semantic_seg") is not None:
            tgt = torch.as_tensor(sample["ori_semantic_seg"][0]).long().to(dev).unsqueeze(0)  # [1,H,W] class ids
            tgt = torch.where((tgt < 0) | (tgt >= self.num_seg), torch.full_like(tgt, -1), tgt)
        else:
            h, w = sample["net_input"]["patch_images"].shape[-2:]
            ids = sample["target"][:, :-1].reshape(-1, h, w).to(dev)
            tgt = class_targets(ids, self.seg_id_offset, self.num_seg, self.padding_idx)
        out = {}

        def put(suffix, scores, target):
            _, ai, ap, al, au = segmentation_metrics(scores, target, hp, wp)
            out.update({f"area_intersect{suffix}": ai, f"area_pred_label{suffix}": ap, f"area_label{suffix}": al,
                        f"area_union{suffix}": au})

        put("", logits, tgt)
        tgt_low = None
        low = sample.get("downsampled_target")
        if low is not None:  # [B, P+1] dictionary ids on the patch grid; the last slot is eos (masked, :279-281)
            assert tuple(low.shape) == tuple(logits.shape[:-1])
            tgt_low = class_targets(low[:, :-1].reshape(-1, hp, wp).to(dev), self.seg_id_offset, self.num_seg, self.padding_idx)
            put("_lowres", logits, tgt_low)  # target grid == patch grid: the x1 "upsample" is the identity
        post = extra.get("resnet_postprocess_probability")
        if post is not None:  # :329-336: the same metric on the propagated probabilities
            put("_resnet_postprocess", post, tgt)
        # "just for display" (:338-341) everywhere except the supervised training branch, where this IS the loss
        ce = (lambda lg, t: PixelCrossEntropyFunction.apply(lg, t, hp, wp, self.eps)) if logits.requires_grad else (
            lambda lg, t: pixel_cross_entropy(lg, t, hp, wp, self.eps))
        loss = ce(logits, tgt if (self.upscale_lprobs or tgt_low is None) else tgt_low)
        out["nll_loss"] = loss
        return loss, out

    # -- seg_criterion.py:246-267 (forward value; 32/512 generalised to patch_image_size) ------
    def imfree_loss_value(self, model, sample):
        S = model.cfg.patch_image_size
        with torch.no_grad():
            _, extra = model(aux_input=sample["aux_input"])
            logits = extra["aux_output"][0]
            ids = sample["text2seg_target"][:, :-1].reshape(-1, S, S).to(logits.device)
            tgt = class_targets(ids, self.seg_id_offset, self.num_seg, self.padding_idx)
            return pixel_cross_entropy(logits, tgt, S // 16, S // 16, self.eps)

    # -- seg_criterion.py:415-588 ---------------------------------------------------------------
    @classmethod
    def reduce_metrics(cls, logging_outputs) -> None:
        """Aggregate the logging outputs of all data-parallel workers: scalar losses averaged by sample_size, the area
        histograms summed per class, aAcc / mIoU / mAcc derived from the summed areas for every metric family."""
        def total(key):
            return sum(log.get(key, 0) for log in logging_outputs)

        sample_size, ntokens = total("sample_size"), total("ntokens")
        metrics.log_scalar("loss", total("loss") / sample_size, sample_size, round=3)
        for key in ("imfree_loss", "seg_loss", "nll_loss"):
            metrics.log_scalar(key, total(key) / sample_size, ntokens, round=3)
        metrics.log_derived("ppl", lambda meters: utils.get_perplexity(meters["nll_loss"].avg))
        for key, val in (("ntokens", ntokens), ("nsentences", total("nsentences")), ("sample_size", sample_size)):
            metrics.log_scalar(key, val, 1, round=3)

        def scalar(v):
            return round(v if isinstance(v, float) else v.item(), 4)

        for sfx in _METRIC_FAMILIES:
            if "area_intersect" + sfx not in logging_outputs[0]:
                continue
            for part in ("intersect", "pred_label", "label", "union"):
                metrics.log_scalar_sum(f"_area_{part}{sfx}", total(f"area_{part}{sfx}"), 1)
            # default arguments bind the suffix of THIS family (the derived meters are evaluated later)
            metrics.log_derived("aAcc" + sfx, lambda m, s=sfx: scalar(
                m["_area_intersect" + s].sum.sum() / m["_area_pred_label" + s].sum.sum()))
            metrics.log_derived("mIoU" + sfx, lambda m, s=sfx: scalar(
                torch.nanmean(m["_area_intersect" + s].sum / m["_area_union" + s].sum)))
            metrics.log_derived("mAcc" + sfx, lambda m, s=sfx: scalar(
                torch.nanmean(m["_area_intersect" + s].sum / m["_area_label" + s].sum)))
        n_total = utils.item(total("total"))
        if n_total > 0:  # report_accuracy family (:574-588); never produced by this criterion's forward
            metrics.log_scalar("total", n_total)
            metrics.log_scalar("n_correct", utils.item(total("n_correct")))
            metrics.log_derived("accuracy", lambda m: round(m["n_correct"].sum * 100.0 / m["total"].sum, 3)
                                if m["total"].sum > 0 else float("nan"))

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:  # seg_criterion.py:590-597
        return True
