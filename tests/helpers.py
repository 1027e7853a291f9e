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


def make_task(num_seg, names=None):
    """Stand-in for SegmentationTask: target_dictionary / tgt_dict with encode_line, cfg, and a bpe whose encode maps a
    word to a deterministic string of 1-3 'subword' ids (what GPT2BPE.encode returns: space-separated id strings)."""
    import types
    import zlib

    from ifseg_b200.fairseq_compat import StubDictionary

    class _Dict(StubDictionary):
        def encode_line(self, line, add_if_not_exist=False, append_eos=False):
            return torch.tensor([4 + int(t) % 50000 for t in line.split()], dtype=torch.int32)

    class _Bpe:
        def encode(self, x):
            h = zlib.crc32(x.encode())
            return " ".join(str((h >> (8 * i)) & 0xFFFF) for i in range(1 + h % 3))

    d = _Dict(num_seg)
    names = names or [f"thing{i} part" if i % 4 == 0 else f"thing{i}" for i in range(num_seg)]
    return types.SimpleNamespace(target_dictionary=d, tgt_dict=d, bpe=_Bpe(),
                                 cfg=types.SimpleNamespace(num_seg_tokens=num_seg, category_list=",".join(names)))
