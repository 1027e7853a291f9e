"""`seg_criterion`, B200-native pieces.  Mirrors criterions/seg_criterion.py of alinlab/ifseg:

  upsample_logits + compute_metric (:237-244, 349-362)  -> one fused kernel (sgf_upsample_argmax): the
      [B,C,H,W] logits the reference materialises (138 MB / image at C=150) never exist;
  compute_imfree_loss / the display CE (:246-267, 340)  -> sgf_upsample_ce_loss (forward value);
  _lazy_initialization (:373-407)                       -> sgf_embedding_bag_mean;
  SegCriterion.forward (:165-235)                       -> same signature and logging_output keys.

Round-1 state: the evaluation branch (model.eval()) incl. the ResNet-feature label propagation
(`resnet_iters > 0`, seg_criterion.py:197-213 -> ops.label_propagation); the training branch
(unsupervised_segmentation: image-free loss + no-grad real-image metrics) returns a loss whose .backward()
runs the hand-written adjoint kernels (PixelCrossEntropyFunction -> train_engine.ImFreeBranchFunction).
"""
import math

import torch

from . import ops
from .segofa import str_bool


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


class SegCriterion:
    """Same constructor arguments / forward contract as the reference criterion.  `task` only needs
    `target_dictionary.index("<seg_0>")`, `.cfg.num_seg_tokens` and `.cfg.category_list` (and `.bpe`,
    `.tgt_dict.encode_line` for init_seg_with_text)."""

    def __init__(self, task, sentence_avg=False, label_smoothing=0.0, upscale_lprobs="true",
                 unsupervised_segmentation="true", criterion_update_freq=1, full_context_alignment="false",
                 init_seg_with_text="true", resnet_topk=3, resnet_prob_temperature=1.0, resnet_iters=0, **unused):
        self.task = task
        self.sentence_avg = sentence_avg
        self.eps = label_smoothing
        self.upscale_lprobs = str_bool(upscale_lprobs)
        self.unsupervised_segmentation = str_bool(unsupervised_segmentation)
        self.full_context_alignment = str_bool(full_context_alignment)
        self.init_seg_with_text = str_bool(init_seg_with_text)
        self.resnet_topk, self.resnet_prob_temperature, self.resnet_iters = resnet_topk, resnet_prob_temperature, resnet_iters
        self.criterion_update_freq = criterion_update_freq
        self.iter = -1
        self.effective_iter = -1
        self.padding_idx = task.target_dictionary.pad()
        self.seg_id_offset = task.target_dictionary.index("<seg_0>")
        self.num_seg = task.cfg.num_seg_tokens
        self.id2rawtext = [x.strip() for x in task.cfg.category_list.split(",")]
        assert len(self.id2rawtext) == self.num_seg

    __call__ = lambda self, *a, **k: self.forward(*a, **k)  # noqa: E731

    # -- seg_criterion.py:373-407 -----------------------------------------------------------
    def lazy_initialization(self, model, class_token_ids):
        """seg_embed_tokens (and the tied seg_projection) <- mean token embedding of each class name.
        class_token_ids: list of 1-D LongTensors (the BPE ids of ' <name>'), one per class."""
        dev = model.encoder.embed_tokens.weight.device
        lens = torch.tensor([len(t) for t in class_token_ids], dtype=torch.long)
        tokens = torch.cat(class_token_ids).to(dev).unsqueeze(0)
        avg = ops.embedding_bag_mean(tokens, lens.cumsum(0).to(dev), model.encoder.embed_tokens.weight.detach(), len(lens))
        avg = avg.to(model.encoder.seg_embed_tokens.weight.dtype)
        model.encoder.seg_embed_tokens.weight.data = avg
        model.decoder.seg_embed_tokens.weight.data = avg
        if not model.decoder.tie_seg_projection:
            model.decoder.seg_projection.weight.data = avg.clone()
        model.invalidate_engine()

    # -- seg_criterion.py:165-235 -------------------------------------------------------------
    def forward(self, model, sample, update_num=0, reduce=True, ema_model=None):
        self.iter += 1
        self.effective_iter = self.iter // self.criterion_update_freq
        if model.training and not self.unsupervised_segmentation:
            raise NotImplementedError("segofa_b200 trains the image-free branch only (--unsupervised-segmentation=true, "
                                      "every shipped recipe); the supervised real-image loss has no backward")
        if model.training:  # seg_criterion.py:178-186
            S = model.cfg.patch_image_size
            net_output = model(full_context_alignment=self.full_context_alignment, aux_input=sample["aux_input"])
            logits = net_output[1]["aux_output"][0]
            ids = sample["text2seg_target"][:, :-1].reshape(-1, S, S).to(logits.device)
            tgt = class_targets(ids, self.seg_id_offset, self.num_seg, self.padding_idx)
            imfree_loss = PixelCrossEntropyFunction.apply(logits, tgt, S // 16, S // 16, self.eps)
            loss = imfree_loss
            with torch.no_grad():
                seg_logits, seg_extra = model(**sample["net_input"], full_context_alignment=self.full_context_alignment)
                seg_loss, metrics = self.compute_loss(seg_logits, seg_extra, sample)
            sample_size = sample["target"].size(0) if self.sentence_avg else 1
            logging_output = {"loss": loss.data, "imfree_loss": imfree_loss.data, "seg_loss": seg_loss.data,
                              "ntokens": sample["ntokens"], "nsentences": sample["nsentences"], "sample_size": sample_size}
            logging_output.update(metrics)
            return loss, sample_size, logging_output
        with torch.no_grad():
            logits, extra = model(**sample["net_input"], full_context_alignment=self.full_context_alignment)
            if self.resnet_iters > 0:  # seg_criterion.py:197-213: label propagation over ResNet-feature neighbours
                feats = extra["encoder_returns"]["image_embed_before_proj"][0]
                prob, _ = ops.label_propagation(feats.contiguous(), logits.float().contiguous(), self.resnet_topk,
                                                self.resnet_iters, self.resnet_prob_temperature)
                eos = prob.new_zeros((prob.shape[0], 1, prob.shape[2]))  # fake eos row (:211)
                extra["resnet_postprocess_probability"] = torch.cat([prob, eos], dim=1)
            seg_loss, metrics = self.compute_loss(logits, extra, sample)
        imfree_loss = torch.zeros(1, device=logits.device)
        loss = seg_loss
        sample_size = sample["target"].size(0) if self.sentence_avg else 1
        logging_output = {"loss": loss.data, "imfree_loss": imfree_loss.data, "seg_loss": seg_loss.data,
                          "ntokens": sample["ntokens"], "nsentences": sample["nsentences"], "sample_size": sample_size}
        logging_output.update(metrics)
        return loss, sample_size, logging_output

    # -- seg_criterion.py:269-347 (evaluation: original-resolution ground truth, batch 1) ------
    def compute_loss(self, logits, extra, sample):
        hp, wp = extra["encoder_returns"]["image_embed_shape"][0]
        dev = logits.device
        if sample.get("ori_semantic_seg") is not None:
            tgt = torch.as_tensor(sample["ori_semantic_seg"][0]).long().to(dev)  # [H,W] class ids
            tgt = tgt.unsqueeze(0)
            tgt = torch.where((tgt < 0) | (tgt >= self.num_seg), torch.full_like(tgt, -1), tgt)
        else:
            h, w = sample["net_input"]["patch_images"].shape[-2:]
            ids = sample["target"][:, :-1].reshape(-1, h, w).to(dev)
            tgt = class_targets(ids, self.seg_id_offset, self.num_seg, self.padding_idx)
        _, ai, ap, al, au = segmentation_metrics(logits, tgt, hp, wp)
        metrics = {"area_intersect": ai, "area_pred_label": ap, "area_label": al, "area_union": au}
        post = extra.get("resnet_postprocess_probability")
        if post is not None:  # seg_criterion.py:329-336: the same metric on the propagated probabilities
            _, ai2, ap2, al2, au2 = segmentation_metrics(post, tgt, hp, wp)
            metrics.update(area_intersect_resnet_postprocess=ai2, area_pred_label_resnet_postprocess=ap2,
                           area_label_resnet_postprocess=al2, area_union_resnet_postprocess=au2)
        loss = pixel_cross_entropy(logits, tgt, hp, wp, self.eps)  # "just for display" (:340)
        metrics["nll_loss"] = loss
        return loss, metrics

    # -- seg_criterion.py:246-267 (forward value; 32/512 generalised to patch_image_size) ------
    def imfree_loss_value(self, model, sample):
        S = model.cfg.patch_image_size
        with torch.no_grad():
            _, extra = model(aux_input=sample["aux_input"])
            logits = extra["aux_output"][0]
            ids = sample["text2seg_target"][:, :-1].reshape(-1, S, S).to(logits.device)
            tgt = class_targets(ids, self.seg_id_offset, self.num_seg, self.padding_idx)
            return pixel_cross_entropy(logits, tgt, S // 16, S // 16, self.eps)

    @staticmethod
    def logging_outputs_can_be_summed() -> bool:  # seg_criterion.py:590-597
        return True
