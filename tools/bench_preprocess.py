"""Input-pipeline measurement: sgf_image_prep_u8 + sgf_segmap_prep_u8 on a B200 against the host
libraries the reference's dataset calls (cv2 + torchvision, one thread as with --num-workers=0).
    python tools/bench_preprocess.py            # prints one JSON line per geometry"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from ifseg_b200 import ops  # noqa: E402
from ifseg_b200.preprocess import IMAGENET_DEFAULT_MEAN as MEAN, IMAGENET_DEFAULT_STD as STD, rescale_size  # noqa: E402


def host_reference(image, seg, C, S, rs_wh, crop, flip, reps):
    import cv2
    from torchvision import transforms

    cv2.setNumThreads(1)
    torch.set_num_threads(1)
    norm = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=MEAN, std=STD)])
    down = transforms.Resize((S // 16, S // 16), transforms.InterpolationMode.NEAREST)
    t0 = time.perf_counter()
    for _ in range(reps):
        s = seg.copy()
        s[s == 0] = 255
        s = s - 1
        s[s == 254] = C
        img = cv2.resize(image[:, :, ::-1].copy(), rs_wh, interpolation=cv2.INTER_LINEAR)
        gt = cv2.resize(s, rs_wh, interpolation=cv2.INTER_NEAREST)
        if crop is not None:
            y, x, h, w = crop
            img, gt = img[y:y + h, x:x + w], gt[y:y + h, x:x + w]
        if flip:
            img, gt = np.flip(img, 1), np.flip(gt, 1)
        t = norm(img[:, :, ::-1].copy())
        g = torch.from_numpy(gt.astype(np.int64))
        d = down(g.unsqueeze(0)).flatten()
        _ = torch.cat([59457 + g.flatten(), torch.tensor([2])]), torch.cat([torch.tensor([0]), 59457 + d]), t
    return (time.perf_counter() - t0) / reps


def main():
    peaks = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))) if os.path.exists(
        os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbps_burst", peaks.get("hbm_gbps", 6550.7)))
    rng = np.random.default_rng(0)
    B = 8
    for name, (h, w), S, C, train in [("coco_val_640x427", (427, 640), 512, 171, False), ("ade_val_683x512", (512, 683), 512, 150, False),
                                      ("coco_train_crop512", (427, 640), 512, 171, True)]:
        images = [rng.integers(0, 256, (h, w, 3), dtype=np.uint8) for _ in range(B)]
        segs = [rng.integers(0, C + 1, (h, w), dtype=np.uint8) for _ in range(B)]
        if train:
            rs_wh, crop, flip = (1150, 767), (100, 300, 512, 512), True
        else:
            rs_wh, crop, flip = rescale_size(w, h, (4 * S, S)), None, False
        rs_hw = (rs_wh[1], rs_wh[0])
        oh, ow = (crop[2], crop[3]) if crop else rs_hw
        pin_i = [torch.from_numpy(x).pin_memory() for x in images]
        pin_s = [torch.from_numpy(x).pin_memory() for x in segs]
        dev_i = [x.cuda() for x in pin_i]
        dev_s = [x.cuda() for x in pin_s]
        out = torch.empty((B, 3, oh, ow), dtype=torch.float32, device="cuda")

        def device_only():
            for b in range(B):
                ops.image_prep_u8(dev_i[b], rs_hw, crop=crop, flip=flip, mean=MEAN, std=STD, out=out[b])
                ops.segmap_prep_u8(dev_s[b], C, rs_hw, (S // 16, S // 16), crop=crop, flip=flip, want_ori=not train,
                                   want_downsampled=train)

        def with_h2d():
            for b in range(B):
                di, ds = pin_i[b].cuda(non_blocking=True), pin_s[b].cuda(non_blocking=True)
                ops.image_prep_u8(di, rs_hw, crop=crop, flip=flip, mean=MEAN, std=STD, out=out[b])
                ops.segmap_prep_u8(ds, C, rs_hw, (S // 16, S // 16), crop=crop, flip=flip, want_ori=not train,
                                   want_downsampled=train)

        res = {}
        for tag, fn in (("device", device_only), ("h2d+device", with_h2d)):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[tag] = e0.elapsed_time(e1) / 20 / B  # ms per image
        # the image kernel alone, for the roofline: bytes = u8 source read once + fp32 CHW written once
        for _ in range(3):
            ops.image_prep_u8(dev_i[0], rs_hw, crop=crop, flip=flip, mean=MEAN, std=STD, out=out[0])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200):
            ops.image_prep_u8(dev_i[i % B], rs_hw, crop=crop, flip=flip, mean=MEAN, std=STD, out=out[i % B])
        e1.record()
        torch.cuda.synchronize()
        k_us = e0.elapsed_time(e1) / 200 * 1e3
        src_bytes = 3 * h * w if not crop else 3 * h * w * (oh * ow) / (rs_hw[0] * rs_hw[1])
        alg = src_bytes + 12 * oh * ow
        host_ms = host_reference(images[0], segs[0], C, S, rs_wh, crop, flip, reps=10) * 1e3
        print(json.dumps(dict(case=name, src_hw=(h, w), resized_hw=rs_hw, out_hw=(oh, ow), batch=B,
                              device_ms_per_image=round(res["device"], 4), h2d_device_ms_per_image=round(res["h2d+device"], 4),
                              device_img_per_s=round(1e3 / res["device"]), h2d_device_img_per_s=round(1e3 / res["h2d+device"]),
                              image_kernel_us=round(k_us, 2), image_kernel_alg_bytes=int(alg),
                              image_kernel_gbps=round(alg / k_us / 1e3, 1), hbm_peak_gbps=hbm,
                              host_cv2_torchvision_ms_per_image=round(host_ms, 2), host_threads=1,
                              speedup_vs_host=round(host_ms / res["h2d+device"], 1))))


if __name__ == "__main__":
    main()
