"""Real-image samples prepared on the device -- the host-side mirror of SegmentationDataset.__getitem__ + collate for
the real-image fields (data/mm_data/segmentation_dataset.py:40-130, 210-301).

The reference decodes the JPEG / label PNG on the host and then resizes, crops, flips, normalises and builds the token
targets per sample on one CPU thread (--num-workers=0).  Here the decoded uint8 arrays are copied to the GPU as bytes
and `sgf_image_prep_u8` / `sgf_segmap_prep_u8` produce the same tensors, bit for bit (tests/test_preprocess_gpu.py
against the cv2 / torchvision golden vectors).  Decoding itself (PIL) and PhotoMetricDistortion stay with the caller.
CUDA-only: there is no host fallback -- the ops raise on CPU tensors."""
import random
from typing import Dict, Optional, Sequence, Tuple

import torch

from . import ops

IMAGENET_DEFAULT_MEAN = (0.485, 0.456, 0.406)
IMAGENET_DEFAULT_STD = (0.229, 0.224, 0.225)


def rescale_size(w: int, h: int, scale: Tuple[int, int]) -> Tuple[int, int]:
    """mmcv.rescale_size for a (long edge, short edge) scale (mmseg Resize(keep_ratio=True)): new (w, h)."""
    long_edge, short_edge = max(scale), min(scale)
    factor = min(long_edge / max(h, w), short_edge / min(h, w))
    return int(w * float(factor) + 0.5), int(h * float(factor) + 0.5)


def random_train_geometry(w: int, h: int, patch_image_size: int, rng: random.Random, ratio_range=(0.5, 2.0),
                          flip_prob=0.5):
    """The random choices of the training transform (:158-163) as plain numbers: Resize(img_scale=(4S, S),
    ratio_range, min_size=S) -> size after resize; RandomCrop(S, S) -> window; RandomFlip(0.5) -> flag.  mmseg draws
    them from numpy's global generator, which cannot be reproduced stream-for-stream; the distributions are the same.
    (RandomCrop's cat_max_ratio re-draw needs the label histogram of the window and is left to the caller.)"""
    S = patch_image_size
    ratio = rng.random() * (ratio_range[1] - ratio_range[0]) + ratio_range[0]  # Resize.random_sample_ratio
    scale = (int(4 * S * ratio), int(S * ratio))
    new_short = max(min(scale), S)  # min_size: the short edge never drops under the crop size (Resize._resize_img)
    new_h, new_w = (new_short * h / w, new_short) if h > w else (new_short, new_short * w / h)
    rs_w, rs_h = rescale_size(w, h, (new_h, new_w))
    cy = rng.randint(0, max(rs_h - S, 0))
    cx = rng.randint(0, max(rs_w - S, 0))
    return (rs_h, rs_w), (cy, cx, min(S, rs_h), min(S, rs_w)), rng.random() < flip_prob


class RealImagePipeline:
    """prepare(image_u8, seg_u8) -> the example dict of __getitem__ (:282-292) with device tensors; collate(examples)
    -> the batch dict of collate (:112-130) for the real-image fields."""

    def __init__(self, num_seg: int, patch_image_size: int, src_item: torch.Tensor, split: str = "valid",
                 mean: Sequence[float] = IMAGENET_DEFAULT_MEAN, std: Sequence[float] = IMAGENET_DEFAULT_STD,
                 seg_id_offset: int = 59457, bos: int = 0, eos: int = 2, pad: int = 1, device="cuda"):
        self.num_seg, self.S = num_seg, patch_image_size
        self.src_item = src_item.long()  # bos + prompt + class names + eos (:267-276), built by the task's BPE
        self.split = split
        self.mean, self.std = tuple(mean), tuple(std)
        self.seg_id_offset, self.bos, self.eos, self.pad = seg_id_offset, bos, eos, pad
        self.device = torch.device(device)

    def prepare(self, image_u8, seg_u8, uniq_id=0, geometry=None) -> Dict:
        """image_u8: uint8 [H, W, 3] RGB (np.asarray(PIL image)); seg_u8: uint8 [H, W] raw label PNG.  geometry =
        ((rs_h, rs_w), (cy, cx, h, w), flip) for the training transform; None = the validation keep-ratio resize."""
        dev = self.device
        img = torch.as_tensor(image_u8).to(dev, non_blocking=True)
        seg = torch.as_tensor(seg_u8).to(dev, non_blocking=True)
        H, W = img.shape[:2]
        S, g = self.S, self.S // 16
        if geometry is None:
            rs_w, rs_h = rescale_size(W, H, (4 * S, S))
            crop, flip = None, False
        else:
            (rs_h, rs_w), crop, flip = geometry
        patch_image = ops.image_prep_u8(img.contiguous(), (rs_h, rs_w), crop=crop, flip=flip, mean=self.mean, std=self.std)
        train = self.split == "train"
        target, prev, down, ori = ops.segmap_prep_u8(seg.contiguous(), self.num_seg, (rs_h, rs_w), (g, g), crop=crop, flip=flip,
                                                     seg_id_offset=self.seg_id_offset, bos_id=self.bos, eos_id=self.eos,
                                                     want_downsampled=train, want_ori=True)
        return {"id": uniq_id, "source": self.src_item, "patch_image": patch_image,
                "patch_mask": torch.tensor([True]), "target": target, "downsampled_target": down,
                "prev_output_tokens": prev, "ori_shape": (H, W, 3), "ori_semantic_seg": ori}

    def collate(self, examples) -> Optional[Dict]:
        if not examples:
            return {}
        dev = self.device

        def merge(key):  # data_utils.collate_tokens: right-padded with pad
            rows = [e[key].to(dev) for e in examples]
            n = max(r.numel() for r in rows)
            out = torch.full((len(rows), n), self.pad, dtype=torch.long, device=dev)
            for i, r in enumerate(rows):
                out[i, :r.numel()] = r
            return out

        src_tokens = merge("source")
        target = merge("target")
        batch = {
            "id": [e["id"] for e in examples],
            "nsentences": len(examples),
            "ntokens": int(sum(e["target"].numel() for e in examples)),
            "net_input": {
                "src_tokens": src_tokens,
                "src_lengths": src_tokens.ne(self.pad).sum(1),
                "patch_images": torch.stack([e["patch_image"] for e in examples], 0),
                "patch_masks": torch.cat([e["patch_mask"] for e in examples]).to(dev),
                "prev_output_tokens": merge("prev_output_tokens"),
            },
            "target": target,
            "ori_shape": [e["ori_shape"] for e in examples],
            "ori_semantic_seg": [e["ori_semantic_seg"] for e in examples],
        }
        if examples[0].get("downsampled_target") is not None:
            batch["downsampled_target"] = merge("downsampled_target")
        return batch
