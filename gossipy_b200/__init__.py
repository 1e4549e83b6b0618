"""gossipy_b200 -- a Blackwell (B200, sm_100a) native gossip-learning engine.

The public surface mirrors ``gossipy`` (reference: ``gossipy/__init__.py:19-34``) so that
``import gossipy_b200 as gossipy`` is a drop-in switch, but the implementation is organised
around *flat parameter rows resident in HBM* instead of deep-copied Python object graphs:

* a "model in flight" is a :class:`~gossipy_b200.engine.arena.Snapshot` (one row of an arena of
  parameter vectors plus its age), not a deep copy of a handler;
* merge / update / evaluate are single fused kernels over rows (``gossipy_b200.ops``);
* ranks (one per GPU) replicate the cheap host-side event schedule and execute only the
  device work of the nodes they own; peers' rows are read directly over NVLink.

This module holds the process-wide runtime objects (reference layer L0).
"""
from __future__ import annotations

import logging
import os
import random
from abc import ABC, abstractmethod
from typing import Any, Dict, Iterable, Optional, Tuple

# NOTE on CUDA lazy loading: kernels of this package may spin on flags written by other GPUs, and
# loading a not-yet-used kernel can need a device-wide synchronisation that would wait behind such a
# spinning kernel.  The extension therefore force-loads ITS OWN kernels on first use (ops/native.py).
# (Setting CUDA_MODULE_LOADING=EAGER globally is not an option: it makes the first cuBLAS call load
# the whole library, minutes on a cold box.)

import numpy as np
import torch

__version__ = "0.1.0"

__all__ = ["LOG", "CACHE", "set_seed", "CacheKey", "CacheItem", "Sizeable", "Cache",
           "GlobalSettings"]


# --------------------------------------------------------------------------------------
# logging (reference: gossipy/__init__.py:94-115 -- a "rich" logger that drops duplicates)
# --------------------------------------------------------------------------------------
class _OncePerMessage(logging.Filter):
    """Lets every distinct message text through exactly once."""

    def __init__(self) -> None:
        super().__init__()
        self._seen = set()

    def filter(self, record: logging.LogRecord) -> bool:  # noqa: A003
        text = record.getMessage()
        if text in self._seen:
            return False
        self._seen.add(text)
        return True


def _make_logger() -> logging.Logger:
    logger = logging.getLogger("gossipy_b200")
    if not logger.handlers:
        try:
            from rich.logging import RichHandler
            handler: logging.Handler = RichHandler(show_path=False)
            handler.setFormatter(logging.Formatter("%(message)s"))
        except Exception:  # rich is optional
            handler = logging.StreamHandler()
            handler.setFormatter(logging.Formatter("[%(levelname)s] %(message)s"))
        logger.addHandler(handler)
        logger.setLevel(logging.INFO)
        logger.propagate = False
        logger.addFilter(_OncePerMessage())
    return logger


LOG = _make_logger()


# --------------------------------------------------------------------------------------
# global settings: device + topology of the job
# --------------------------------------------------------------------------------------
class _SingletonMeta(type):
    _made: Dict[type, Any] = {}

    def __call__(cls, *a, **kw):
        inst = _SingletonMeta._made.get(cls)
        if inst is None:
            inst = super().__call__(*a, **kw)
            _SingletonMeta._made[cls] = inst
        return inst


class GlobalSettings(metaclass=_SingletonMeta):
    """Process-wide settings (reference: ``gossipy/__init__.py:46-91``).

    Beyond the reference's single torch device this also carries the *job topology*: the
    rank / world size of the one-process-per-GPU job and the backend used for the hot ops
    (``"cuda"`` = hand written sm_100a kernels, ``"torch"`` = plain PyTorch, used on CPU).
    Unlike the reference the device is looked up lazily by handlers, so it may be changed
    between experiments.
    """

    def __init__(self) -> None:
        self._device = torch.device("cpu")
        self.rank = 0
        self.world_size = 1
        self.reference_compat = False  # mimic selected reference quirks (see docs/QUIRKS.md)
        self._allow_tf32 = False
        # replay the forward + backward of generic (autograd) models from a CUDA graph per handler and batch shape
        # instead of launching every layer's kernels from Python (model/handler.py: _graph_fwd_bwd)
        self.cuda_graphs = os.environ.get("GOSSIPY_CUDA_GRAPHS", "1") != "0"
        # rows of convolutional models keep their filters in channels-last order and feed NHWC batches (cuDNN's tensor-core
        # kernels run without layout conversions): "auto" = handlers created while the device is a GPU, True = always
        # (also on CPU: tests), False = never
        self.channels_last = {"0": False, "1": True}.get(os.environ.get("GOSSIPY_CHANNELS_LAST", ""), "auto")

    # -- numerics -----------------------------------------------------------------------
    @property
    def allow_tf32(self) -> bool:
        """``False`` (default): the fused GPU kernels compute like the reference's fp32 -- tensor-core products are
        error compensated (3xTF32).  ``True``: plain tf32 products (operands truncated to 10 mantissa bits, the
        analogue of ``torch.backends.cuda.matmul.allow_tf32``), ~25 % faster per local update."""
        return self._allow_tf32

    @allow_tf32.setter
    def allow_tf32(self, on: bool) -> None:
        self._allow_tf32 = bool(on)
        from . import ops
        ops.set_train_impl("tc8-tf32" if on else "")
        ops.set_eval_tf32(bool(on))

    # -- device -------------------------------------------------------------------------
    def auto_device(self) -> torch.device:
        self._device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        return self._device

    def set_device(self, device_name: str) -> torch.device:
        if device_name == "auto":
            return self.auto_device()
        self._device = torch.device(device_name)
        return self._device

    def get_device(self) -> torch.device:
        return self._device

    # -- topology -----------------------------------------------------------------------
    def set_topology(self, rank: int, world_size: int) -> None:
        assert 0 <= rank < world_size
        self.rank, self.world_size = rank, world_size

    def is_cuda(self) -> bool:
        return self._device.type == "cuda"


def set_seed(seed: int = 0) -> None:
    """Seed every RNG the engine draws from (reference: ``gossipy/__init__.py:118-131``).

    Additionally seeds the CUDA generators and the engine's counter-based device RNG, which
    keys every device-side draw by ``(seed, node, purpose, counter)`` so that a run is
    reproducible regardless of how nodes are placed on GPUs.
    """
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    from .engine import rng as _rng
    _rng.set_base_seed(seed)


# --------------------------------------------------------------------------------------
# size accounting
# --------------------------------------------------------------------------------------
class Sizeable(ABC):
    """Anything whose size in "atoms" (scalars) can be reported (ref ``__init__.py:134-156``)."""

    @abstractmethod
    def get_size(self) -> int:
        ...


def atoms_of(value: Any) -> int:
    """Number of atoms of a payload element; shared by Message / CacheItem accounting."""
    if value is None:
        return 0
    if isinstance(value, (bool, int, float, np.integer, np.floating)):
        return 1
    if isinstance(value, Sizeable):
        return int(value.get_size())
    raise TypeError("Cannot compute the size of %r" % (value,))


# --------------------------------------------------------------------------------------
# the in-flight model store.  reference: gossipy/__init__.py:159-380
# --------------------------------------------------------------------------------------
class CacheKey(Sizeable):
    """Hashable handle of an in-flight model: ``(owner, age, ...)``."""

    __slots__ = ("key",)

    def __init__(self, *args: Any) -> None:
        self.key: Tuple[Any, ...] = tuple(args)

    def get(self) -> Tuple[Any, ...]:
        return self.key

    def get_size(self) -> int:
        val = CACHE[self]
        if val is None:
            return 0
        try:
            return atoms_of(val)
        except TypeError:
            LOG.warning("Impossible to compute the size of %s. Set to 0." % val)
            return 0

    def __hash__(self) -> int:
        return hash(self.key)

    def __eq__(self, other: Any) -> bool:
        return isinstance(other, CacheKey) and self.key == other.key

    def __ne__(self, other: Any) -> bool:
        return not self == other

    def __repr__(self) -> str:
        return str(self.key)


class CacheItem(Sizeable):
    """A cached value with a reference count (reference ``__init__.py:200-280``)."""

    __slots__ = ("_value", "_refs")

    def __init__(self, value: Any) -> None:
        self._value = value
        self._refs = 1
        self._publish()

    def add_ref(self) -> None:
        self._refs += 1
        self._publish()

    def del_ref(self) -> Any:
        self._refs -= 1
        self._publish()
        return self._value

    def _publish(self) -> None:
        # snapshots consult this count so that `release()` by one receiver does not free a
        # row that other in-flight messages still reference
        try:
            self._value._cache_refs = self._refs
        except AttributeError:
            pass

    def is_referenced(self) -> bool:
        return self._refs > 0

    def get(self) -> Any:
        return self._value

    def get_size(self) -> int:
        v = self._value
        if isinstance(v, (tuple, list)):
            total = 0
            for el in v:
                try:
                    total += atoms_of(el)
                except TypeError:
                    LOG.warning("Impossible to compute the size of %s. Set to 0." % el)
            return max(total, 1)
        try:
            return atoms_of(v)
        except TypeError:
            LOG.warning("Impossible to compute the size of %s. Set to 0." % v)
            return 0

    def __repr__(self) -> str:
        return repr(self._value)

    def __str__(self) -> str:
        return "CacheItem(%s)" % (self._value,)


class Cache:
    """Reference-counted store of models that are "on the wire".

    Semantics follow the reference (``gossipy/__init__.py:283-377``): pushing an existing key
    only bumps its refcount, popping decrements and frees at zero.  Values are usually
    :class:`~gossipy_b200.engine.arena.Snapshot` objects whose parameter row is returned to the
    arena when the entry dies (``release()`` hook), which fixes the reference's leak of
    dropped messages without changing the API.
    """

    def __init__(self) -> None:
        # one store per Cache *instance* (the reference shares a class attribute)
        self._cache: Dict[CacheKey, CacheItem] = {}

    def push(self, key: CacheKey, value: Any) -> None:
        item = self._cache.get(key)
        if item is None:
            self._cache[key] = CacheItem(value)
        else:
            item.add_ref()
            _release(value)  # the duplicate snapshot is not needed

    def pop(self, key: CacheKey) -> Any:
        item = self._cache.get(key)
        if item is None:
            return None
        value = item.del_ref()
        if not item.is_referenced():
            del self._cache[key]
        return value

    def drop(self, key: CacheKey) -> None:
        """Pop and release: used for messages that are lost (drop / offline receiver)."""
        item = self._cache.get(key)
        if item is None:
            return
        value = item.del_ref()
        if not item.is_referenced():
            del self._cache[key]
            _release(value)

    def clear(self) -> None:
        for item in self._cache.values():
            value = item.get()
            try:  # force: the entries die with the cache whatever their reference counts say
                value._cache_refs = 0
            except AttributeError:
                pass
            _release(value)
        self._cache.clear()

    def __getitem__(self, key: CacheKey) -> Any:
        item = self._cache.get(key)
        return None if item is None else item.get()

    def __contains__(self, key: CacheKey) -> bool:
        return key in self._cache

    def load(self, cache_dict: Dict[CacheKey, Any]) -> None:
        self._cache = cache_dict

    def get_cache(self) -> Dict[CacheKey, Any]:
        return self._cache

    def __len__(self) -> int:
        return len(self._cache)

    def __repr__(self) -> str:
        return str(self._cache)


def _release(value: Any) -> None:
    rel = getattr(value, "release", None)
    if callable(rel):
        rel()


CACHE = Cache()
"""The global store of in-flight models (reference: ``gossipy/__init__.py:380``)."""
