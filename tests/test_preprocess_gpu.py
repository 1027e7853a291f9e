"""-m gpu: sgf_image_prep_u8 / sgf_segmap_prep_u8 (real-image input pipeline on the device) against
the cv2 + torchvision golden vectors and the numpy oracle -- bit-exact (integer / byte work, and IEEE fp32 for the
normalisation)."""
import numpy as np
import pytest
import torch

from test_preprocess_cpu import FIELDS, golden_cases

pytestmark = pytest.mark.gpu


def _run(image, seg, C, S, rs_wh, crop, flip):
    from ifseg_b200 import ops
    from oracle.preprocess import IMAGENET_DEFAULT_MEAN as MEAN, IMAGENET_DEFAULT_STD as STD

    rs_hw = (rs_wh[1], rs_wh[0])
    img = torch.from_numpy(image).cuda()
    sg = torch.from_numpy(seg).cuda()
    out = {"patch_image": ops.image_prep_u8(img, rs_hw, crop=crop, flip=flip, mean=MEAN, std=STD)}
    out["target"], out["prev_output_tokens"], out["downsampled_target"], out["ori_semantic_seg"] = ops.segmap_prep_u8(
        sg, C, rs_hw, (S // 16, S // 16), crop=crop, flip=flip, want_downsampled=True, want_ori=True)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def test_golden_vectors_bit_exact(cuda_device):
    n = 0
    for name, z, m in golden_cases():
        out = _run(z[name + "/image"], z[name + "/seg"], m["C"], m["S"], m["rs_wh"], m["crop"], m["flip"])
        for k in FIELDS:
            assert np.array_equal(out[k], z[f"{name}/{k}"]), (name, k, int((out[k] != z[f"{name}/{k}"]).sum()))
        n += 1
    assert n == 7


@pytest.mark.parametrize("h,w,S,C", [(427, 640, 512, 171), (640, 427, 480, 150), (375, 500, 480, 15), (1024, 2048, 512, 150),
                                     (960, 1280, 480, 15)])
def test_full_size_images_vs_oracle(cuda_device, h, w, S, C):
    """Validation geometry at dataset-like sizes (COCO 640x427, ADE up to 2048 wide; 960x1280 -> 480x640 is the exact-2x
    INTER_AREA shortcut), and a training geometry (random scale, crop window, flip) on the same image."""
    import random

    from ifseg_b200.preprocess import random_train_geometry, rescale_size
    from oracle import preprocess as P

    rng = np.random.default_rng(h * 7 + w)
    image = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    seg = rng.integers(0, C + 2, (h // 9 + 1, w // 9 + 1), dtype=np.uint8).repeat(9, 0).repeat(9, 1)[:h, :w].copy()
    seg[seg == C + 1] = 255
    geos = [(rescale_size(w, h, (4 * S, S)), None, False)]
    (rs_h, rs_w), crop, flip = random_train_geometry(w, h, S, random.Random(h + w))
    geos.append(((rs_w, rs_h), crop, True))
    for rs_wh, crop, flip in geos:
        ref = P.prepare(image, seg, C, S, rs_wh=rs_wh, crop=crop, flip=flip)
        out = _run(image, seg, C, S, rs_wh, crop, flip)
        for k in FIELDS:
            assert np.array_equal(out[k], ref[k]), (k, rs_wh, crop, flip, int((out[k] != ref[k]).sum()))


def test_pipeline_sample_feeds_model_and_criterion(cuda_device):
    """RealImagePipeline.prepare + collate build the sample dict the criterion's validation branch takes
    (seg_criterion.py:194-215, 283-287): ori-resolution metrics on a non-square keep-ratio grid."""
    import types

    from helpers import build_cuda_model, load_prompts
    from ifseg_b200.fairseq_compat import StubDictionary
    from ifseg_b200.preprocess import RealImagePipeline
    from ifseg_b200.seg_criterion import SegCriterion

    C, S = 15, 128
    model, _ = build_cuda_model("segofa_base", C, S, seed=0)
    prompt = torch.tensor(load_prompts()[str(C)], dtype=torch.long)
    pipe = RealImagePipeline(C, S, prompt, split="valid")
    rng = np.random.default_rng(2)
    image = rng.integers(0, 256, (150, 200, 3), dtype=np.uint8)
    seg = rng.integers(0, C + 1, (150, 200), dtype=np.uint8)
    ex = pipe.prepare(image, seg, uniq_id=7)
    assert ex["patch_image"].shape == (3, 128, 171) and ex["prev_output_tokens"].numel() == 65
    sample = pipe.collate([ex])
    assert sample["net_input"]["patch_images"].shape == (1, 3, 128, 171)
    task = types.SimpleNamespace(target_dictionary=StubDictionary(C),
                                 cfg=types.SimpleNamespace(num_seg_tokens=C, category_list=",".join(f"c{i}" for i in range(C))))
    crit = SegCriterion(task, init_seg_with_text="false")
    model.eval()
    loss, sample_size, log = crit(model, sample)
    assert torch.isfinite(loss).all()
    # every labelled original-resolution pixel is counted once
    assert int(log["area_label"].sum().item()) == int((seg != 0).sum() - (seg == 255).sum())
