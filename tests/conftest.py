import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _restore_reference_after_install():
    """install() patches the reference package process-wide; every test leaves the reference as it found it (the live
    differential suites compare against the UNPATCHED reference)."""
    yield
    mod = sys.modules.get("speechclip_plus_b200.install")
    if mod is not None:
        mod.uninstall()


def load_golden(name: str):
    """Load tests/golden/<name>.npz as a dict of torch tensors / python scalars."""
    out = {}
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False) as z:
        for k in z.files:
            a = z[k]
            if a.dtype.kind in "US":
                out[k] = str(a)
            elif a.ndim == 0 and a.dtype == np.bool_:
                out[k] = bool(a)
            else:
                out[k] = torch.from_numpy(np.array(a))
    return out


def golden_names(prefix: str):
    return sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.startswith(prefix) and f.endswith(".npz"))


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """max |a-b| / max |b| -- the 'relative error' used for every float comparison in this suite."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.abs().max().item()
    if denom == 0.0:
        return (a - b).abs().max().item()
    return (a - b).abs().max().item() / denom


def norm_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """||a-b||_2 / ||b||_2."""
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    denom = b.norm().item()
    return (a - b).norm().item() / (denom if denom > 0 else 1.0)
