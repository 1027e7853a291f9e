"""One optimizer update of the shipped IFSeg recipe on the B200-native engines -- the sequence
custom_fairseq's Trainer.train_step runs (trainer.py:745-1049):

  SegCriterion.forward (seg_criterion.py:165-193)
      image-free branch forward + compute_imfree_loss + backward        -> SegOFATrainEngine.forward_backward
      inference-mode real-image forward (:185) + compute_loss metrics   -> SegOFAEngine (live operands) + sgf_upsample_argmax/ce
  gradient all-reduce (DDP, C1)                                         -> bucketed NCCL all-reduce of the flat fp32
                                                                           gradient arena, launched per layer bucket as
                                                                           the backward reaches it (overlaps the backward)
  multiply_grads(world / sample_size), clip_grad_norm (:879-895)        -> sgf_sumsq + device-side clip coefficient
  FP16Optimizer.step: fp32 master Adam + copy back (:108-222)           -> sgf_adam_step on the flat arena + bf16 operand refresh

Nothing in the step synchronises with the host.
"""
from typing import Dict, Optional

import torch

from . import ops
from .seg_criterion import class_targets
from .train_engine import SegOFATrainEngine


class SegOFATrainer:
    def __init__(self, model, lr=5e-5, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.1, clip_norm=1.0,
                 label_smoothing=0.0, seg_id_offset=59457, process_group=None, eval_real_image=True, stochastic=True,
                 seed=1, supervised=False):
        self.model = model
        self.supervised = supervised  # --unsupervised-segmentation=false: the real-image loss trains (seg_criterion.py:188-192)
        self.engine = SegOFATrainEngine(model, stochastic=stochastic, seed=seed)  # dropout / DropPath of the recipe
        self.lr, self.betas, self.eps, self.weight_decay, self.clip_norm = lr, betas, eps, weight_decay, clip_norm
        self.label_smoothing = label_smoothing
        self.seg_id_offset = seg_id_offset
        self.eval_real_image = eval_real_image
        self.world = 1
        self.pg = process_group
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.world = torch.distributed.get_world_size(process_group)
        self.engine.grad_sync = self._sync_bucket if self.world > 1 else None
        self._works = []
        self._side = None
        self.num_updates = 0

    # gradient exchange: called by the engine as soon as a contiguous gradient range is final
    def _sync_bucket(self, lo, hi):
        if hi > lo:
            self._works.append(torch.distributed.all_reduce(self.engine.arena.grad32[lo:hi], group=self.pg, async_op=True))

    def to_device(self, sample):
        """_prepare_sample (trainer.py:1257-1295): host -> device copies of one collated sample."""
        dev = self.engine.device

        def mv(x):
            if torch.is_tensor(x):
                return x.to(dev, non_blocking=True)
            if isinstance(x, dict):
                return {k: mv(v) for k, v in x.items()}
            return x

        return mv(sample)

    def train_step(self, sample, check_pads=True) -> Dict[str, torch.Tensor]:
        cfg = self.model.cfg
        eng = self.engine
        S = cfg.patch_image_size
        C = cfg.num_seg
        B = sample["net_input"]["src_tokens"].shape[0]
        dev = eng.device
        self._works = []
        if self.supervised:
            ni = sample["net_input"]
            c = eng.forward_train(ni, check_pads=check_pads)
            h, w = ni["patch_images"].shape[-2:]
            rt = class_targets(sample["target"][:, :-1].reshape(B, h, w).to(dev), self.seg_id_offset, C, cfg.padding_idx)
            seg_loss, dlogits = eng.loss_and_dlogits(c, rt, self.label_smoothing)
            with torch.no_grad():  # display metrics from the same logits (compute_loss, :301-305)
                _, areas = ops.upsample_argmax(c["logits"], c["h"], c["w"], h, w, target=rt)
            eng.backward_from(c, dlogits)
            for wk in self._works:
                wk.wait()
            gnorm = eng.optimizer_step(self.lr, self.betas, self.eps, self.weight_decay, self.clip_norm,
                                       grad_mult=1.0 / self.world)
            self.num_updates += 1
            return {"loss": seg_loss, "seg_loss": seg_loss, "imfree_loss": torch.zeros_like(seg_loss), "gnorm": gnorm,
                    "area_intersect": areas[0], "area_pred_label": areas[1], "area_label": areas[2],
                    "area_union": areas[1] + areas[2] - areas[0]}
        ids = sample["text2seg_target"][:, :-1].reshape(B, S, S).to(dev)
        tgt = class_targets(ids, self.seg_id_offset, C, cfg.padding_idx)
        c = eng.forward_train(sample["aux_input"], check_pads=check_pads)
        imfree_loss, dlogits = eng.loss_and_dlogits(c, tgt, self.label_smoothing)
        log = {"loss": imfree_loss, "imfree_loss": imfree_loss}
        main = torch.cuda.current_stream()
        if self.eval_real_image:
            # `with torch.inference_mode(): model(**net_input)` + compute_loss (display metrics, seg_criterion.py:185):
            # independent of the backward, so it runs on a side stream underneath it (its small stem kernels fill
            # the SMs the backward's wave tails leave idle); joined before the optimizer touches the weights
            if self._side is None:
                self._side = torch.cuda.Stream()
            self._side.wait_stream(main)
            ni = sample["net_input"]
            # a padded prompt takes the key-padding path (host check as encoder_module.py:742; skipped under graph capture)
            ni_pads = check_pads and bool(ni["src_tokens"].eq(cfg.padding_idx).any())
            with torch.cuda.stream(self._side), torch.no_grad():
                enc = eng.inf.encode(ni["src_tokens"], patch_images=ni["patch_images"], patch_masks=ni["patch_masks"],
                                     has_pads=ni_pads)
                logits, _ = eng.inf.decode(enc, ni["prev_output_tokens"])
                hp, wp = enc["hw"]
                h, w = ni["patch_images"].shape[-2:]
                rt = class_targets(sample["target"][:, :-1].reshape(B, h, w).to(dev), self.seg_id_offset, C, cfg.padding_idx)
                _, areas = ops.upsample_argmax(logits, hp, wp, h, w, target=rt)
                seg_loss, _ = ops.upsample_ce_loss(logits, rt, hp, wp, self.label_smoothing)
                union = areas[1] + areas[2] - areas[0]
                for t in (seg_loss, areas, union):
                    t.record_stream(main)
            log.update(seg_loss=seg_loss, area_intersect=areas[0], area_pred_label=areas[1], area_label=areas[2],
                       area_union=union)
        eng.backward_from(c, dlogits)
        if self.eval_real_image:
            main.wait_stream(self._side)
        for wk in self._works:
            wk.wait()
        gnorm = eng.optimizer_step(self.lr, self.betas, self.eps, self.weight_decay, self.clip_norm,
                                   grad_mult=1.0 / self.world)
        self.num_updates += 1
        log["gnorm"] = gnorm
        return log


class TrainSession:
    """Static-shape training loop: the sample lives in fixed device buffers, one train_step is captured in a
    CUDA graph (single-GPU) and replayed; `load` refreshes the buffers from (pinned) host memory."""

    def __init__(self, trainer: SegOFATrainer, sample_host, use_cuda_graph=True, warmup=2):
        self.trainer = trainer
        self.stream = torch.cuda.Stream()
        self.static = trainer.to_device(sample_host)
        torch.cuda.synchronize()
        self.graph = None
        self.launches_per_step = 0
        with torch.cuda.stream(self.stream):
            for _ in range(max(1, warmup)):  # eager: builds caches (position bias, index tensors, Adam moments)
                n0 = ops.launch_count()
                self.out = trainer.train_step(self.static, check_pads=False)
                self.launches_per_step = ops.launch_count() - n0
        self.stream.synchronize()
        import os

        # world > 1: capturing the bucketed NCCL all-reduces works (c10d records them as cross-stream dependencies, so
        # they still overlap the backward: 31.1 vs 31.7 ms eager at 2 GPUs) but process-group teardown then hangs
        # with this torch/NCCL build, so multi-GPU sessions run eagerly unless SGF_GRAPH_NCCL=1
        if use_cuda_graph and (trainer.world == 1 or os.environ.get("SGF_GRAPH_NCCL") == "1"):
            try:
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=self.stream):
                    self.out = trainer.train_step(self.static, check_pads=False)
                self.graph = graph
            except Exception as e:  # noqa: BLE001 -- e.g. a c10d build that refuses capture: stay eager
                if trainer.world == 1:
                    raise
                import warnings

                warnings.warn(f"CUDA-graph capture of the multi-GPU train step failed ({e}); running eagerly")
                torch.cuda.synchronize()
                self.graph = None

    def load(self, sample_host):
        """host (pinned) sample -> the static device buffers (async on the session stream)."""
        def cp(dst, src):
            if torch.is_tensor(dst):
                dst.copy_(src, non_blocking=True)
            elif isinstance(dst, dict):
                for k in dst:
                    cp(dst[k], src[k])

        with torch.cuda.stream(self.stream):
            cp(self.static, sample_host)

    def step(self):
        with torch.cuda.stream(self.stream):
            if self.graph is not None:
                self.graph.replay()
            else:
                self.out = self.trainer.train_step(self.static, check_pads=False)
        return self.out
