"""Shared test helpers (oracle access lives here: tests may import oracle/, the product may not)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLD, f"golden_{name}.pt"), map_location="cpu")


def load_prompts():
    return json.load(open(os.path.join(GOLD, "prompts.json")))


def oracle_cfg(cfg):
    from oracle import restated as R

    return R.SegOFAConfig(**{k: getattr(cfg, k) for k in R.SegOFAConfig.__dataclass_fields__ if hasattr(cfg, k)})


def rel_l2(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-20)).item()


def build_cuda_model(arch, num_seg, size, seed=0):
    from ifseg_b200.segofa import SegOFAModel
    from ifseg_b200.synthetic import generate_state_dict

    model = SegOFAModel.from_config(arch, num_seg, size)
    sd = generate_state_dict(model.cfg, seed)
    model.load_state_dict(sd, strict=True)
    return model.cuda().eval(), sd
