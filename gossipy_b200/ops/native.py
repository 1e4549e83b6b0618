"""Loader of the in-tree sm_100a extension ``gossipy_b200._C``."""
from __future__ import annotations

import importlib

_mod = None
_err = None


def _try_import():
    global _mod, _err
    if _mod is None and _err is None:
        try:
            _mod = importlib.import_module("gossipy_b200._C")
        except Exception as exc:  # noqa: BLE001
            _err = exc
    return _mod


def native_available() -> bool:
    return _try_import() is not None


def require_native() -> None:
    if _try_import() is None:
        raise RuntimeError(
            "gossipy_b200: CUDA tensors were passed but the sm_100a extension gossipy_b200._C is "
            "not built/loadable (%r). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "at the repo root. There is deliberately no eager fallback on GPU." % (_err,))


_preloaded = set()


def native():
    """The extension module; on first use per CUDA device every kernel is force-loaded (CUDA loads
    functions lazily, and loading one while another kernel spins on a cross-GPU flag can deadlock)."""
    require_native()
    import torch
    if torch.cuda.is_available():
        dev = torch.cuda.current_device()
        if dev not in _preloaded:
            _preloaded.add(dev)
            _mod.preload()
    return _mod
