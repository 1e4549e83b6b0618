"""Flat parameter rows: the memory layout every fused kernel works on.

A model's learnable state is ONE contiguous fp32 vector ("row"):

    [ parameters in ``parameters()`` order | float buffers (e.g. BN running stats) | pad ]

padded to a multiple of 32 floats (128 B) so that rows in an arena are 128-byte aligned and
kernels can use 128-bit accesses end to end.  Integer buffers (``num_batches_tracked``) are kept
in a small side vector and merged with ``max`` (policy for SURVEY B12).  ``bind`` re-points a
module's tensors at views of a row, so autograd / ``module(x)`` keep working while merge,
optimizer step and snapshot are single launches over the row.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

ROW_ALIGN = 32  # floats


class FlatLayout:
    """Offsets of every parameter / float buffer of a module inside a flat row."""

    def __init__(self, module: nn.Module, channels_last: bool = False) -> None:
        # channels_last: 4-D parameters (convolution filters [O, I, kh, kw]) are stored in the row as [O, kh, kw, I] and
        # bound as permuted views, i.e. as ``torch.channels_last`` tensors -- cuDNN then runs its NHWC tensor-core
        # kernels without converting filters and activations back and forth on every call.  Which element sits at
        # which row index is a private matter of the layout: merges, snapshots and optimizers are element-wise.
        self.channels_last = bool(channels_last) and any(p.dim() == 4 for p in module.parameters())
        self.param_names: List[str] = []
        self.entries: List[Tuple[str, torch.Size, int, int]] = []  # name, shape, offset, numel
        off = 0
        # channels-last rows: every tensor starts on a 32-byte boundary (cuDNN's NHWC kernels use 128-bit accesses on the
        # filter / gradient pointers; a filter that starts at an odd float offset faults with "misaligned address").
        # The pads are ordinary row elements that stay zero (zero gradient); plain rows stay densely packed -- the
        # fused MLP / logreg kernels address [W1 | b1 | W2 | b2] directly.
        align = 8 if self.channels_last else 1
        for name, p in module.named_parameters():
            off = (off + align - 1) // align * align
            self.param_names.append(name)
            self.entries.append((name, p.shape, off, p.numel()))
            off += p.numel()
        self.n_params = off  # == model.get_size() for ordinary models (channels-last rows: plus alignment pads)
        self.float_buffers: List[Tuple[str, torch.Size, int, int]] = []
        self.int_buffers: List[str] = []
        for name, b in module.named_buffers():
            if b is None:
                continue
            if b.dtype.is_floating_point:
                off = (off + align - 1) // align * align
                self.float_buffers.append((name, b.shape, off, b.numel()))
                off += b.numel()
            else:
                self.int_buffers.append(name)
        self.numel = off
        self.padded = max(ROW_ALIGN, (off + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN)
        self.requires_grad = [bool(p.requires_grad) for p in module.parameters()]

    def _view(self, flat: torch.Tensor, shape: torch.Size) -> torch.Tensor:
        """The tensor of ``shape`` stored in the 1-D row slice ``flat``."""
        if self.__dict__.get("channels_last", False) and len(shape) == 4:
            o, i, kh, kw = shape
            return flat.view(o, kh, kw, i).permute(0, 3, 1, 2)
        return flat.view(shape)

    # -- moving data between a module and a row ------------------------------------------
    @torch.no_grad()
    def gather(self, module: nn.Module, row: torch.Tensor) -> None:
        """Copy the module's current values into ``row``."""
        sd = dict(module.named_parameters())
        sd.update(dict(module.named_buffers()))
        if self.__dict__.get("channels_last", False):
            row[:self.numel].zero_()            # alignment pads between the tensors
        for name, shape, off, n in self.entries + self.float_buffers:
            self._view(row[off:off + n], shape).copy_(sd[name].detach())
        if self.padded > self.numel:
            row[self.numel:self.padded].zero_()

    @torch.no_grad()
    def bind(self, module: nn.Module, row: torch.Tensor,
             grad_row: Optional[torch.Tensor] = None) -> None:
        """Make the module's parameters / float buffers views of ``row`` (and grads of ``grad_row``)."""
        params = dict(module.named_parameters())
        for name, shape, off, n in self.entries:
            p = params[name]
            p.data = self._view(row[off:off + n], shape)
            if grad_row is not None and p.requires_grad:
                p.grad = self._view(grad_row[off:off + n], shape)
        if self.float_buffers:
            owners = _buffer_owners(module)
            for name, shape, off, n in self.float_buffers:
                mod, leaf = owners[name]
                mod._buffers[leaf] = row[off:off + n].view(shape)

    def views(self, row: torch.Tensor) -> Dict[str, torch.Tensor]:
        """``{name: view}`` for ``torch.func.functional_call`` on a snapshot row."""
        return {name: self._view(row[off:off + n], shape)
                for name, shape, off, n in self.entries + self.float_buffers}

    # -- index helpers used by partitioned / sampled merges -------------------------------
    def param_slices(self) -> List[Tuple[int, torch.Size]]:
        return [(off, shape) for _, shape, off, _ in self.entries]


def _buffer_owners(module: nn.Module) -> Dict[str, Tuple[nn.Module, str]]:
    out = {}
    for mod_name, mod in module.named_modules():
        for leaf in mod._buffers:
            full = leaf if not mod_name else mod_name + "." + leaf
            out[full] = (mod, leaf)
    return out


def int_buffer_state(module: nn.Module, names: List[str]) -> Dict[str, torch.Tensor]:
    bufs = dict(module.named_buffers())
    return {n: bufs[n] for n in names}
