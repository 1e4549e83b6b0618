"""Collectives over the gossip nodes' parameter rows.

:class:`SymmetricAllReduce` is the all-to-all round of decentralised SGD on a clique (every node
averages all models, ``gossipy/simul.py:756-852`` with uniform mixing) as ONE self-synchronising
kernel per GPU (``csrc/kernels/nvls.cu``):

* every rank owns a *symmetric* buffer (``torch.distributed._symmetric_memory`` is used for the
  plumbing only: allocation, handle exchange, multicast binding);
* with NVLink SHARP the kernel reduces inside the NVSwitch (``multimem.ld_reduce``), otherwise it
  pulls the peers' buffers with P2P loads;
* flags in the buffers' tails order the ranks (ready / done epochs) -- no NCCL call, no host sync.

On CPU (gloo plumbing runs) it degrades to ``dist.all_reduce``.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import runtime as prt

_FLAG_WORDS = 64      # room for ready[16] + done[16] (+ padding to keep the row 128-byte aligned)


class SymmetricAllReduce:
    """Mean of one ``numel``-float vector per rank.  Collective construction."""

    def __init__(self, numel: int, device: torch.device, use_multicast: Optional[bool] = None) -> None:
        self.numel = int(numel)
        assert self.numel % 4 == 0, "rows are padded to a multiple of 32 floats"
        self.device = torch.device(device)
        self.world = prt.world() if prt.active() else 1
        self.rank = prt.rank() if prt.active() else 0
        self.epoch = 0
        self.mc_ptr = 0
        self.kind = "local"
        if self.device.type != "cuda":
            self.buf = torch.zeros(self.numel, dtype=torch.float32)
            self.kind = "gloo" if self.world > 1 else "local"
            return
        if self.world == 1:
            self.buf = torch.zeros(self.numel + _FLAG_WORDS, dtype=torch.float32, device=self.device)
            self._bufs = [self.buf.data_ptr()]
            self._flags = [self.buf.data_ptr() + 4 * self.numel]
            self.kind = "p2p"
            return
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm
        self.buf = symm.empty(self.numel + _FLAG_WORDS, dtype=torch.float32, device=self.device)
        self.buf.zero_()
        torch.cuda.synchronize(self.device)
        self._hdl = symm.rendezvous(self.buf, dist.group.WORLD.group_name)
        self._bufs = [int(p) for p in self._hdl.buffer_ptrs]
        self._flags = [p + 4 * self.numel for p in self._bufs]
        mc = 0
        try:
            mc = int(self._hdl.multicast_ptr)
        except Exception:
            mc = 0
        if use_multicast is False:
            mc = 0
        if use_multicast is True and mc == 0:
            raise RuntimeError("NVLS multicast was requested but is not available on this system")
        self.mc_ptr = mc
        self.kind = "nvls" if mc else "p2p"
        dist.barrier()

    @property
    def contribution(self) -> torch.Tensor:
        """This rank's input vector (write it on the stream that later calls :meth:`mean_into`)."""
        return self.buf[:self.numel]

    def mean_into(self, out: torch.Tensor, n_total: int) -> None:
        """``out = (sum over ranks of contribution) / n_total`` on the current stream."""
        self.epoch += 1
        if self.kind in ("nvls", "p2p"):
            from ..ops import _count
            from ..ops.native import native
            native().allreduce_mean(out, self.mc_ptr, self._bufs, self._flags, self.rank, self.epoch,
                                    1.0 / float(n_total), self.numel)
            _count()
            return
        if self.kind == "gloo":
            import torch.distributed as dist
            tmp = self.buf.clone()
            dist.all_reduce(tmp)
            out[:self.numel].copy_(tmp / float(n_total))
        else:
            out[:self.numel].copy_(self.buf / float(n_total))
