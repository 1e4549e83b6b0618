"""Test configuration: the ``gpu`` marker, reference import shims and state isolation."""
import os
import sys
import types

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with `pytest -m gpu` on a B200)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _fresh_state():
    import gossipy_b200 as g
    from gossipy_b200.engine import arena
    g.CACHE.clear()
    g.GlobalSettings().reference_compat = False
    g.GlobalSettings().set_device("cpu")
    g.set_seed(1234)
    yield
    g.CACHE.clear()
    g.GlobalSettings().set_device("cpu")


def _import_reference():
    """Import the read-only reference with the three environment shims of SURVEY §0."""
    ref_root = "/root/reference"
    if not os.path.isdir(os.path.join(ref_root, "gossipy")):
        return None
    for name in ("matplotlib", "matplotlib.pyplot", "pyparsing"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mod = types.ModuleType(name)
                if name == "pyparsing":
                    mod.ParseSyntaxException = Exception
                sys.modules[name] = mod
    if "matplotlib" in sys.modules and not hasattr(sys.modules["matplotlib"], "pyplot"):
        sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    if ref_root not in sys.path:
        sys.path.append(ref_root)
    try:
        import gossipy  # noqa: F401
        import gossipy.model.handler as H
        _auc = H.roc_auc_score
        if not getattr(H, "_b200_patched", False):
            H.roc_auc_score = lambda *a, **k: np.float64(_auc(*a, **k))
            H._b200_patched = True
        import gossipy.simul as S

        class _It:
            def __init__(self, it): self.it = it
            def __iter__(self): return iter(self.it)
            def close(self): pass
        S.track = lambda it, description="": _It(it)
        return gossipy
    except Exception:
        return None


@pytest.fixture(scope="session")
def ref():
    mod = _import_reference()
    if mod is None:
        pytest.skip("reference not importable")
    return mod
