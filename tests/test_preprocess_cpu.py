"""-m "not gpu": the numpy restatement of the real-image input pipeline (oracle/preprocess.py) against the golden
vectors produced by cv2 + torchvision through the reference's call sequence (oracle/make_golden_preprocess.py), and the
host-side geometry rules of ifseg_b200/preprocess.py."""
import os
import random

import numpy as np
import pytest

from helpers import GOLD

GOLDEN = os.path.join(GOLD, "golden_preprocess.npz")
FIELDS = ("patch_image", "target", "prev_output_tokens", "downsampled_target", "ori_semantic_seg")


def golden_cases():
    z = np.load(GOLDEN)
    for name in sorted({k.split("/")[0] for k in z.files}):
        S, C, rw, rh, cy, cx, ch, cw, flip = (int(v) for v in z[name + "/meta"])
        yield name, z, dict(S=S, C=C, rs_wh=(rw, rh), crop=(cy, cx, ch, cw), flip=bool(flip))


def test_oracle_matches_cv2_and_torchvision_golden_vectors():
    from oracle import preprocess as P

    n = 0
    for name, z, m in golden_cases():
        out = P.prepare(z[name + "/image"], z[name + "/seg"], m["C"], m["S"], rs_wh=m["rs_wh"], crop=m["crop"], flip=m["flip"])
        for k in FIELDS:
            assert out[k].dtype == z[f"{name}/{k}"].dtype or k != "patch_image"
            assert np.array_equal(out[k], z[f"{name}/{k}"]), (name, k)  # bit-exact, fp32 image included
        n += 1
    assert n == 7


def test_validation_size_rule_and_identity_cases():
    from oracle import preprocess as P
    from ifseg_b200 import preprocess as H

    for (w, h, S) in [(640, 427, 512), (427, 640, 512), (500, 375, 480), (2048, 300, 512), (512, 512, 512), (100, 75, 64)]:
        assert H.rescale_size(w, h, (4 * S, S)) == P.rescale_size(w, h, (4 * S, S))
        nw, nh = H.rescale_size(w, h, (4 * S, S))
        assert min(nw, nh) == S or max(nw, nh) == 4 * S  # one of the two edges hits its bound
    for name, z, m in golden_cases():
        if name.startswith("val"):
            H_, W_ = z[name + "/image"].shape[:2]
            assert H.rescale_size(W_, H_, (4 * m["S"], m["S"])) == m["rs_wh"]
    # unchanged size -> cv2 returns the input: the fixed-point arithmetic must be the identity
    img = np.random.default_rng(1).integers(0, 256, (17, 23, 3), dtype=np.uint8)
    assert np.array_equal(P.resize_linear_u8(img, 23, 17), img)
    assert np.array_equal(P.resize_nearest_u8(img, 23, 17), img)


def test_cv2_itself_when_importable():
    """Extra pin at full size against the library itself (skipped where opencv is absent)."""
    cv2 = pytest.importorskip("cv2")
    from oracle import preprocess as P

    rng = np.random.default_rng(5)
    for (h, w, dh, dw) in [(427, 640, 512, 767), (375, 500, 480, 640), (600, 800, 300, 400), (333, 500, 512, 769)]:
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        assert np.array_equal(P.resize_linear_u8(img, dw, dh), cv2.resize(img, (dw, dh), interpolation=cv2.INTER_LINEAR))
        seg = rng.integers(0, 256, (h, w), dtype=np.uint8)
        assert np.array_equal(P.resize_nearest_u8(seg, dw, dh), cv2.resize(seg, (dw, dh), interpolation=cv2.INTER_NEAREST))


def test_train_geometry_rules():
    from ifseg_b200.preprocess import random_train_geometry

    rng = random.Random(3)
    flips = 0
    for _ in range(200):
        w, h = rng.randint(200, 900), rng.randint(200, 900)
        (rs_h, rs_w), (cy, cx, ch, cw), flip = random_train_geometry(w, h, 512, rng)
        assert min(rs_h, rs_w) >= 512 and (ch, cw) == (512, 512)  # min_size keeps the crop inside the resized image
        assert 0 <= cy <= rs_h - ch and 0 <= cx <= rs_w - cw
        assert abs(rs_w / rs_h - w / h) < 0.01 * w / h + 2 / min(rs_h, rs_w)  # aspect ratio kept
        assert 512 <= min(rs_h, rs_w) <= 1024  # ratio_range (0.5, 2.0) of the (2048, 512) scale
        flips += flip
    assert 60 < flips < 140


def test_pipeline_has_no_cpu_fallback():
    import torch
    from ifseg_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.image_prep_u8(torch.zeros((4, 4, 3), dtype=torch.uint8), (4, 4), mean=(0, 0, 0), std=(1, 1, 1))
