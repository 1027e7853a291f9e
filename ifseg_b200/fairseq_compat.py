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


# ----------------------------------------------------------------------------------------
# criterion side of the boundary: FairseqCriterion / register_criterion / FairseqDataclass / metrics / utils
# (custom_fairseq/fairseq/criterions/__init__.py:19-26, fairseq_criterion.py:15-62, logging/metrics.py:111-170)
# ----------------------------------------------------------------------------------------
try:  # pragma: no cover - real fairseq deployment
    from fairseq import metrics, utils  # type: ignore
    from fairseq.criterions import CRITERION_REGISTRY, FairseqCriterion, register_criterion  # type: ignore
    from fairseq.dataclass import FairseqDataclass  # type: ignore
except Exception:
    import inspect as _inspect
    import types as _types
    from dataclasses import dataclass as _dataclass

    from torch.nn.modules.loss import _Loss

    CRITERION_REGISTRY = {}

    @_dataclass
    class FairseqDataclass:
        pass

    class FairseqCriterion(_Loss):
        def __init__(self, task):
            super().__init__()
            self.task = task
            if hasattr(task, "target_dictionary"):
                d = task.target_dictionary
                self.padding_idx = d.pad() if d is not None else -100

        @classmethod
        def build_criterion(cls, cfg, task):
            """constructor arguments are looked up by name on the config (fairseq_criterion.py:31-62)"""
            kw = {}
            for prm in _inspect.signature(cls).parameters.values():
                if prm.name == "task":
                    kw["task"] = task
                elif prm.name == "cfg":
                    kw["cfg"] = cfg
                elif hasattr(cfg, prm.name):
                    kw[prm.name] = getattr(cfg, prm.name)
                elif prm.default is prm.empty:
                    raise NotImplementedError(f"cannot infer criterion argument {prm.name}")
            return cls(**kw)

        @staticmethod
        def logging_outputs_can_be_summed() -> bool:
            return False

    def register_criterion(name, dataclass=None):
        def deco(cls):
            if name in CRITERION_REGISTRY:
                raise ValueError(f"Cannot register duplicate criterion ({name})")
            if not issubclass(cls, FairseqCriterion):
                raise ValueError(f"Criterion ({name}: {cls.__name__}) must extend FairseqCriterion")
            cls.__dataclass = dataclass
            CRITERION_REGISTRY[name] = cls
            return cls

        return deco

    class _Avg:
        def __init__(self, round=None):
            self.round, self.sum, self.count = round, 0, 0

        def update(self, val, n=1):
            self.sum = self.sum + val * n
            self.count = self.count + n

        @property
        def avg(self):
            return self.sum / self.count if self.count else self.sum

        @property
        def smoothed_value(self):
            return _round(self.avg, self.round)

    class _Sum:
        def __init__(self, round=None):
            self.round, self.sum = round, 0

        def update(self, val):
            self.sum = self.sum + val

        @property
        def smoothed_value(self):
            return _round(self.sum, self.round)

    def _round(v, nd):
        import torch as _t

        if _t.is_tensor(v):
            if v.numel() != 1:
                return v
            v = v.item()
        return round(v, nd) if nd is not None else v

    class _Metrics:
        """The three logging calls reduce_metrics uses, aggregated into one flat meter dict."""

        def __init__(self):
            self.reset()

        def reset(self):
            self.meters, self.derived = {}, {}

        def log_scalar(self, key, value, weight=1, priority=10, round=None):
            self.meters.setdefault(key, _Avg(round)).update(value, weight)

        def log_scalar_sum(self, key, value, priority=10, round=None):
            self.meters.setdefault(key, _Sum(round)).update(value)

        def log_derived(self, key, fn, priority=20):
            self.derived[key] = fn

        def get_smoothed_values(self, name=None):
            out = {k: m.smoothed_value for k, m in self.meters.items() if not k.startswith("_")}
            out.update({k: fn(self.meters) for k, fn in self.derived.items()})
            return out

    metrics = _Metrics()

    def _get_perplexity(loss, round=2, base=2):
        if loss is None:
            return 0.0
        try:
            return _round(base ** loss, round)
        except OverflowError:
            return float("inf")

    def _item(t):
        return t.item() if hasattr(t, "item") else t

    utils = _types.SimpleNamespace(get_perplexity=_get_perplexity, item=_item)


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
