"""Model base class (behavioural reference: ``gossipy/model/__init__.py:22-74``)."""
from __future__ import annotations

from abc import ABC, abstractmethod

import torch
from torch.nn import ParameterList

from .. import Sizeable

__all__ = ["TorchModel"]


class TorchModel(torch.nn.Module, Sizeable, ABC):
    """A ``torch.nn.Module`` that knows its size and how to initialise itself.

    In the device engine a model's parameters (and float buffers) are *views into one flat
    fp32 row* of a parameter arena (see :mod:`gossipy_b200.engine.flat`), which is what lets
    merge / optimizer / snapshot be single fused kernels.  ``fused_family()`` lets a model
    advertise a hand-written training/eval kernel family (``"mlp1"``, ``"logreg"``, ...).
    """

    def __init__(self, *args, **kwargs) -> None:
        super().__init__()

    @abstractmethod
    def init_weights(self, *args, **kwargs) -> None:
        """(Re-)initialise the weights."""

    def get_size(self) -> int:
        return sum(int(p.numel()) for p in self.parameters())

    def get_params_list(self) -> ParameterList:
        return ParameterList(self.parameters())

    def fused_family(self):
        """``None`` or ``(family_name, dims_tuple)`` when a fused sm_100a kernel family applies."""
        return None

    def __repr__(self) -> str:
        return str(self)

    def __str__(self) -> str:
        return "%s(size=%d)" % (self.__class__.__name__, self.get_size())
