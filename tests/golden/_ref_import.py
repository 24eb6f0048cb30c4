"""Import the reference's own hot-path modules from /root/reference (build container only).

Used ONLY by ``make_golden.py`` (and by the optional differential test that is skipped when
/root/reference is absent).  Nothing here is needed on the GPU box: the fixtures it produces
are committed as ``tests/golden/*.npz``.

The reference's package ``__init__`` files import third-party packages that are not installed
here (clip, fairseq, s3prl, pytorch_lightning, librosa ...).  None of the hot-path arithmetic
lives in them, so they are replaced by empty stub modules before ``avssl`` is imported.
"""
from __future__ import annotations

import importlib
import importlib.machinery
import importlib.util
import os
import sys
import types

REF_ROOT = os.environ.get("SCP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "avssl", "module", "losses.py"))


def load_leaf(rel_path: str, name: str):
    """Load one reference file by path, bypassing the package __init__ files."""
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF_ROOT, rel_path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class _Anything:
    """Attribute sink used to satisfy `from stub import Name` statements."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, item):
        return _Anything()


def _stub(name: str, **attrs):
    mod = types.ModuleType(name)
    mod.__spec__ = importlib.machinery.ModuleSpec(name, loader=None)
    mod.__path__ = []  # behaves as a package so that sub-modules can be stubbed too

    def _getattr(item):
        if item.startswith("__"):
            raise AttributeError(item)
        return _Anything

    mod.__getattr__ = _getattr
    for k, v in attrs.items():
        setattr(mod, k, v)
    sys.modules[name] = mod
    return mod


def import_avssl():
    """Return the imported reference package ``avssl`` with third-party imports stubbed."""
    import torch
    import torch.nn as nn
    import torchvision  # noqa: F401  (import the real one before pandas is stubbed)

    if "avssl" in sys.modules:
        return sys.modules["avssl"]

    class _LightningModule(nn.Module):
        def save_hyperparameters(self, *a, **k):
            pass

        def log(self, *a, **k):
            pass

        def log_dict(self, *a, **k):
            pass

    for name in ["clip", "clip.simple_tokenizer", "fairseq", "fairseq.models", "fairseq.models.hubert",
                 "fairseq.models.hubert.hubert", "fairseq.models.wav2vec", "fairseq.models.wav2vec.wav2vec2",
                 "fairseq.utils", "fairseq.checkpoint_utils", "s3prl", "s3prl.utility", "s3prl.utility.download",
                 "s3prl.hub", "librosa", "editdistance", "sacrebleu", "plotly", "plotly.express",
                 "plotly.graph_objects", "wandb", "soundfile", "matplotlib", "matplotlib.pyplot"]:
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                _stub(name)
    if "pytorch_lightning" not in sys.modules:
        try:
            importlib.import_module("pytorch_lightning")
        except Exception:
            pl = _stub("pytorch_lightning", LightningModule=_LightningModule)
            pl.Trainer = _Anything
            pl.seed_everything = lambda *a, **k: None
            for sub in ["loggers", "callbacks", "loggers.wandb", "utilities", "utilities.distributed",
                        "callbacks.model_checkpoint"]:
                _stub("pytorch_lightning." + sub)
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    return importlib.import_module("avssl")
