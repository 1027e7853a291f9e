"""Inference session: images in -> masks out, the call a user of the reference makes in
visualize_segmentation_web.ipynb cell 4 (model(**net_input) -> logits[:, :-1] -> bilinear
upsample -> argmax), without the ResNet label-propagation / CRF post-processing.

  * the static-shape forward (stem -> encoder -> decoder -> seg_projection -> upsample+argmax)
    is captured once into a CUDA graph (~250 kernel launches replayed with one cudaGraphLaunch);
  * host I/O is pipelined on separate streams over double-buffered pinned staging:
    H2D of batch i+1 and D2H of mask i-1 overlap the compute of batch i.
"""
from typing import Iterable, Iterator, Optional

import torch

from . import ops


class SegmentationSession:
    def __init__(self, model, batch: int, image_size: int, src_tokens, use_cuda_graph: bool = True,
                 out_size: Optional[tuple] = None, mask_dtype=torch.int64):
        p = next(model.parameters())
        if not p.is_cuda:
            raise RuntimeError("SegmentationSession needs the model on a CUDA device (no CPU fallback)")
        self.model = model.eval()
        self.engine = model.engine()
        self.device = p.device
        self.B, self.S = batch, image_size
        self.out_hw = out_size or (image_size, image_size)
        tok = torch.as_tensor(src_tokens, dtype=torch.long)
        if tok.dim() == 1:
            tok = tok.unsqueeze(0).repeat(batch, 1)
        self.has_pads = bool(tok.eq(model.cfg.padding_idx).any())
        self.tokens = tok.to(self.device)
        self.prev = torch.zeros(batch, 1, dtype=torch.long, device=self.device)
        self.images = torch.zeros(batch, 3, image_size, image_size, dtype=torch.float32, device=self.device)
        self.masks = None
        self.logits = None
        self.graph = None
        self.launches_per_step = 0
        self.compute = torch.cuda.Stream(device=self.device)
        self.h2d = torch.cuda.Stream(device=self.device)
        self.d2h = torch.cuda.Stream(device=self.device)
        self._stage_in = [torch.empty_like(self.images) for _ in range(2)]
        self._pin_in = [torch.empty(self.images.shape, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._stage_out = None
        self._pin_out = None
        with torch.no_grad():
            with torch.cuda.stream(self.compute):
                for _ in range(2):  # warm-up: func attributes, shape caches, allocator pools
                    ops.reset_launch_count()
                    self._forward()
                    self.launches_per_step = ops.launch_count()
                self.compute.synchronize()
                if use_cuda_graph:
                    self.graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(self.graph, stream=self.compute):
                        self._forward()
        self._stage_out = [torch.empty_like(self.masks) for _ in range(2)]
        # three pinned result buffers: the one handed to the caller is not written again before the NEXT yield
        self._pin_out = [torch.empty(self.masks.shape, dtype=self.masks.dtype).pin_memory() for _ in range(3)]
        torch.cuda.synchronize(self.device)

    def _forward(self):
        eng = self.engine
        enc = eng.encode(self.tokens, patch_images=self.images, has_pads=self.has_pads)
        logits, _ = eng.decode(enc, self.prev)
        self.logits = logits
        self.masks = eng.predict_mask(logits, enc["hw"], self.out_hw)

    # -- device-resident step (inputs already in HBM) --------------------------------------
    def step_device(self):
        """One forward->mask pass over self.images on the session's compute stream."""
        with torch.cuda.stream(self.compute):
            if self.graph is not None:
                self.graph.replay()
            else:
                with torch.no_grad():
                    self._forward()
        return self.masks

    # -- host-to-host -----------------------------------------------------------------------
    def infer(self, host_images: torch.Tensor) -> torch.Tensor:
        """Blocking single-batch call: pageable/pinned host images -> host masks."""
        for out in self.infer_stream([host_images]):
            return out.clone()

    def infer_stream(self, batches: Iterable[torch.Tensor], pinned_inputs: bool = False) -> Iterator[torch.Tensor]:
        """Pipelined: yields the host mask tensor of each batch.  The yielded tensor is one of three rotating pinned
        buffers and stays valid until the generator is resumed TWICE more (i.e. across the next yield); clone it to keep
        it longer.  With pinned_inputs=True the given tensors are used as the pinned H2D source directly (no staging
        memcpy on the host)."""
        ev_in = [torch.cuda.Event() for _ in range(2)]
        ev_free = [torch.cuda.Event() for _ in range(2)]
        ev_done = [torch.cuda.Event() for _ in range(2)]
        ev_out = [torch.cuda.Event() for _ in range(3)]
        pending = []
        for i, hb in enumerate(batches):
            s, o = i & 1, i % 3
            if i >= 2:
                ev_out[(i - 2) % 3].synchronize()  # batch i-2 is on the host; its stage buffers (slot s) are free again
                yield pending.pop(0)
            src = hb
            if not (pinned_inputs and hb.is_pinned()):
                self._pin_in[s].copy_(hb)
                src = self._pin_in[s]
            with torch.cuda.stream(self.h2d):
                if i >= 2:
                    self.h2d.wait_event(ev_free[s])
                self._stage_in[s].copy_(src, non_blocking=True)
                ev_in[s].record(self.h2d)
            with torch.cuda.stream(self.compute):
                self.compute.wait_event(ev_in[s])
                self.images.copy_(self._stage_in[s], non_blocking=True)
                ev_free[s].record(self.compute)
                if i >= 2:
                    self.compute.wait_event(ev_out[(i - 2) % 3])  # _stage_out[s] has been drained by the D2H of batch i-2
                self.step_device()
                self._stage_out[s].copy_(self.masks, non_blocking=True)
                ev_done[s].record(self.compute)
            with torch.cuda.stream(self.d2h):
                self.d2h.wait_event(ev_done[s])
                self._pin_out[o].copy_(self._stage_out[s], non_blocking=True)
                ev_out[o].record(self.d2h)
            pending.append(self._pin_out[o])
        for k, out in enumerate(pending):
            self.d2h.synchronize()
            yield out

    @property
    def h2d_bytes_per_step(self):
        return self.images.numel() * self.images.element_size()

    @property
    def d2h_bytes_per_step(self):
        return self.masks.numel() * self.masks.element_size()
