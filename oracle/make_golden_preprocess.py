"""Generates tests/golden/golden_preprocess.npz: outputs of the libraries the reference's dataset calls (cv2,
torchvision) for the call sequence of data/mm_data/segmentation_dataset.py:210-265, on small seeded inputs.
mmseg / mmcv are not installed in this image; their two relevant steps are a size rule (restated) and a direct call
into cv2.resize (mmcv.imresize(..., interpolation='bilinear' | 'nearest', backend='cv2')), which is what runs here.

    python oracle/make_golden_preprocess.py        # needs cv2 + torchvision; run in this container
"""
import os

import cv2
import numpy as np
import torch
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "tests", "golden", "golden_preprocess.npz")
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)  # imagenet_default_mean_and_std=True (:148-150)


def reference_sequence(image_rgb, seg_raw, num_seg, S, rs_wh, crop, flip, seg_id_offset=59457, bos=0, eos=2):
    """The statements of __getitem__ (:216-262) with mmseg's transforms written out as the cv2 calls they make."""
    image_arr = image_rgb[:, :, ::-1].copy()  # to BGR (:219)
    seg = seg_raw.copy()
    seg[seg == 0] = 255
    seg = seg - 1
    seg[seg == 254] = num_seg
    ori = seg.copy()
    img = cv2.resize(image_arr, rs_wh, interpolation=cv2.INTER_LINEAR)       # mmcv.imrescale / imresize 'bilinear'
    gt = cv2.resize(seg, rs_wh, interpolation=cv2.INTER_NEAREST)            # ... 'nearest' for seg_fields
    if crop is not None:                                                     # RandomCrop.crop
        y, x, h, w = crop
        img, gt = img[y:y + h, x:x + w, ...], gt[y:y + h, x:x + w]
    if flip:                                                                 # mmcv.imflip horizontal
        img, gt = np.flip(img, axis=1), np.flip(gt, axis=1)
    img = img[:, :, ::-1].copy()  # to RGB (:238, :255)
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)])
    patch_image = norm(img)
    gt = torch.from_numpy(gt.astype(np.int64))
    down = transforms.Resize((S // 16, S // 16), transforms.InterpolationMode.NEAREST)(gt.unsqueeze(0)).flatten()
    codes = seg_id_offset + down
    return dict(patch_image=patch_image.numpy(), target=torch.cat([seg_id_offset + gt.flatten(), torch.tensor([eos])]).numpy(),
                prev_output_tokens=torch.cat([torch.tensor([bos]), codes]).numpy(),
                downsampled_target=torch.cat([codes, torch.tensor([eos])]).numpy(), ori_semantic_seg=ori.astype(np.int64))


def cases():
    """(name, H, W, S, num_seg, rs_wh or None (validation rule), crop, flip)"""
    return [
        ("val_landscape", 75, 100, 64, 15, None, None, False),     # keep-ratio upscale: 64 x 85
        ("val_portrait", 120, 67, 64, 150, None, None, False),     # keep-ratio downscale
        ("val_exact_half", 128, 192, 64, 15, None, None, False),   # 2x decimation: the INTER_AREA shortcut
        ("val_identity", 64, 96, 64, 15, None, None, False),       # same size
        ("train_crop_flip", 90, 131, 64, 171, (186, 128), (37, 59, 64, 64), True),
        ("train_crop", 53, 47, 64, 15, (91, 103), (11, 3, 64, 64), False),
        ("train_down_crop", 140, 131, 64, 150, (97, 104), (40, 12, 64, 64), True),
    ]


def main():
    from oracle.preprocess import rescale_size

    rng = np.random.default_rng(20261017)
    blob = {}
    for name, H, W, S, C, rs_wh, crop, flip in cases():
        # smooth-ish image with hard edges (exercises the rounding), labels in blobs incl. 0 and 255
        image = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        image[H // 3: H // 2] = np.linspace(0, 255, W, dtype=np.uint8)[None, :, None]
        seg = rng.integers(0, C + 1, (H // 7 + 1, W // 7 + 1), dtype=np.uint8).repeat(7, 0).repeat(7, 1)[:H, :W].copy()
        seg[:3, :5] = 255
        if rs_wh is None:
            rs_wh = rescale_size(W, H, (4 * S, S))
        out = reference_sequence(image, seg, C, S, tuple(int(v) for v in rs_wh), crop, flip)
        blob[f"{name}/image"], blob[f"{name}/seg"] = image, seg
        blob[f"{name}/meta"] = np.array([S, C, rs_wh[0], rs_wh[1], *(crop if crop else (0, 0, rs_wh[1], rs_wh[0])), int(flip)])
        for k, v in out.items():
            blob[f"{name}/{k}"] = v
    np.savez_compressed(OUT, **blob)
    print("wrote", OUT, os.path.getsize(OUT), "bytes; cv2", cv2.__version__)


if __name__ == "__main__":
    import sys

    sys.path.insert(0, os.path.join(HERE, ".."))
    main()
