"""Tracing / profiling helpers (the reference has none beyond a progress bar).

* :func:`nvtx_range` -- NVTX ranges around the simulator phases (visible in Nsight Systems / ncu);
* :class:`DeviceTimer` -- CUDA-event timing of a code region on the current stream, reported as the
  MAX over ranks (BASELINE.json: every multi-GPU number is timed on the device as the max over ranks);
* :class:`PhaseProfile` -- a ``SimulationEventReceiver`` that records per-round device time and the
  number of native kernel launches.
"""
from __future__ import annotations

import contextlib
import time
from typing import Dict, List, Optional

import torch

ENABLE_NVTX = False


@contextlib.contextmanager
def nvtx_range(name: str):
    """NVTX range when tracing is enabled and CUDA is present; otherwise free."""
    on = ENABLE_NVTX and torch.cuda.is_available()
    if on:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if on:
            torch.cuda.nvtx.range_pop()


class DeviceTimer:
    """``with DeviceTimer() as t: ...`` then ``t.ms`` (device time on CUDA, wall time on CPU) and
    ``t.max_over_ranks()``."""

    def __init__(self, device: Optional[torch.device] = None) -> None:
        self.cuda = torch.cuda.is_available() and (device is None or torch.device(device).type == "cuda")
        self.ms = 0.0

    def __enter__(self) -> "DeviceTimer":
        if self.cuda:
            self._a, self._b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            self._a.record()
        else:
            self._t0 = time.perf_counter()
        return self

    def __exit__(self, *exc) -> None:
        if self.cuda:
            self._b.record()
            torch.cuda.synchronize()
            self.ms = self._a.elapsed_time(self._b)
        else:
            self.ms = (time.perf_counter() - self._t0) * 1e3

    def max_over_ranks(self) -> float:
        from ..parallel import runtime as prt
        if not prt.active():
            return self.ms
        import torch.distributed as dist
        t = torch.tensor([self.ms], dtype=torch.float64, device="cuda" if self.cuda else "cpu")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())


class PhaseProfile:
    """Attach with ``sim.add_receiver(PhaseProfile())``: wall time between evaluations (= one round,
    the host is synchronised by the metric read-back) and native launches per round."""

    def __init__(self) -> None:
        self.rounds: List[Dict[str, float]] = []
        self._t = time.perf_counter()
        self._launches = self._count()

    @staticmethod
    def _count() -> int:
        from .. import ops
        return ops.launch_count

    def update_message(self, failed, msg=None) -> None:
        pass

    def update_timestep(self, t) -> None:
        pass

    def update_end(self) -> None:
        pass

    def update_evaluation(self, round, on_user, evaluation) -> None:
        now, n = time.perf_counter(), self._count()
        self.rounds.append({"t": round, "ms": (now - self._t) * 1e3, "launches": n - self._launches})
        self._t, self._launches = now, n
