"""Registration boundary.

With the reference's vendored fairseq importable (the normal `--user-dir` deployment) the
model registers itself through fairseq's own decorators and derives from its base classes, so
`train.py` / `trainer.py` / `tasks.build_model` find it exactly like the reference model
(custom_fairseq/fairseq/models/__init__.py:55-194).  Where fairseq cannot be imported (this
repo's CPU/GPU test boxes: no omegaconf/hydra/bitarray), a minimal registry with the same
decorator signatures and base-class constructors keeps the module importable and testable.
"""
import torch.nn as nn

try:  # pragma: no cover - exercised only in a real fairseq deployment
    from fairseq.models import (  # type: ignore
        ARCH_CONFIG_REGISTRY,
        ARCH_MODEL_REGISTRY,
        MODEL_REGISTRY,
        FairseqEncoder,
        FairseqEncoderDecoderModel,
        FairseqIncrementalDecoder,
        register_model,
        register_model_architecture,
    )

    HAVE_FAIRSEQ = True
except Exception:  # ImportError and friends (partial installs fail in many ways)
    HAVE_FAIRSEQ = False
    MODEL_REGISTRY = {}
    ARCH_MODEL_REGISTRY = {}
    ARCH_CONFIG_REGISTRY = {}

    class BaseFairseqModel(nn.Module):
        @classmethod
        def add_args(cls, parser):
            pass

        @classmethod
        def build_model(cls, args, task):
            raise NotImplementedError

        def set_num_updates(self, num_updates):
            for m in self.modules():
                if m is not self and hasattr(m, "set_num_updates"):
                    m.set_num_updates(num_updates)

        def upgrade_state_dict(self, state_dict):
            self.upgrade_state_dict_named(state_dict, "")

        def upgrade_state_dict_named(self, state_dict, name):
            pass

    class FairseqEncoderDecoderModel(BaseFairseqModel):
        def __init__(self, encoder, decoder):
            super().__init__()
            self.encoder = encoder
            self.decoder = decoder

        def max_positions(self):
            return (self.encoder.max_positions(), self.decoder.max_positions())

    class FairseqEncoder(nn.Module):
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

    class FairseqIncrementalDecoder(nn.Module):
        def __init__(self, dictionary):
            super().__init__()
            self.dictionary = dictionary

    def register_model(name, dataclass=None):
        def deco(cls):
            if name in MODEL_REGISTRY:
                raise ValueError(f"Cannot register duplicate model ({name})")
            if not issubclass(cls, BaseFairseqModel):
                raise ValueError(f"Model ({name}: {cls.__name__}) must extend BaseFairseqModel")
            MODEL_REGISTRY[name] = cls
            return cls

        return deco

    def register_model_architecture(model_name, arch_name):
        def deco(fn):
            if model_name not in MODEL_REGISTRY:
                raise ValueError(f"Cannot register model architecture for unknown model type ({model_name})")
            if arch_name in ARCH_MODEL_REGISTRY:
                raise ValueError(f"Cannot register duplicate model architecture ({arch_name})")
            if not callable(fn):
                raise ValueError(f"Model architecture must be callable ({arch_name})")
            ARCH_MODEL_REGISTRY[arch_name] = MODEL_REGISTRY[model_name]
            ARCH_CONFIG_REGISTRY[arch_name] = fn
            return fn

        return deco


class StubDictionary:
    """Stand-in for the task dictionary when no fairseq task exists (bench / tests): the sizes
    and special ids SegmentationTask.setup_task produces (tasks/mm_tasks/segmentation.py:109-136)."""

    def __init__(self, num_seg):
        self.num_seg = num_seg

    def __len__(self):
        return 59457 + self.num_seg + 1

    def __contains__(self, sym):
        return True

    def bos(self):
        return 0

    def pad(self):
        return 1

    def eos(self):
        return 2

    def unk(self):
        return 3

    def index(self, sym):
        return {"<bin_0>": 58457, "<seg_0>": 59457}[sym]
