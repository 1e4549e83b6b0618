"""Token-account flow control (Danner et al. 2018).

Behavioural reference: ``gossipy/flow_control.py:22-236``.  Each strategy is defined by two
pure functions of the token balance ``a`` -- the *proactive* send probability and the
*reactive* message count -- so the same tables drive this Python class hierarchy and the C++
scheduler (``csrc/scheduler.cpp`` takes ``(kind, C, A, k)`` from :meth:`TokenAccount.spec`).
"""
from __future__ import annotations

from abc import ABC, abstractmethod
from typing import Tuple

import numpy as np

__all__ = ["TokenAccount", "PurelyProactiveTokenAccount", "PurelyReactiveTokenAccount",
           "SimpleTokenAccount", "GeneralizedTokenAccount", "RandomizedTokenAccount"]


class TokenAccount(ABC):
    """A per-node account of send tokens."""

    KIND = -1  # id understood by the C++ scheduler

    def __init__(self) -> None:
        self.n_tokens = 0

    def add(self, n: int = 1) -> None:
        self.n_tokens += n

    def sub(self, n: int = 1) -> None:
        self.n_tokens = max(0, self.n_tokens - n)

    @abstractmethod
    def proactive(self) -> float:
        """Probability of sending when the node's timer fires."""

    @abstractmethod
    def reactive(self, utility: int) -> int:
        """Number of messages to send in reaction to a received (useful) message."""

    def spec(self) -> Tuple[int, int, int, int]:
        """``(kind, C, A, k)`` for the native scheduler."""
        return (self.KIND, getattr(self, "capacity", 0), getattr(self, "reactivity", 0),
                getattr(self, "k", 0))

    def __repr__(self) -> str:
        return "%s(tokens=%d)" % (self.__class__.__name__, self.n_tokens)


class PurelyProactiveTokenAccount(TokenAccount):
    """Always send on timeout, never react.  FIX(B21): has a real ``n_tokens`` field."""
    KIND = 0

    def proactive(self) -> float:
        return 1

    def reactive(self, utility: int) -> int:
        return 0


class PurelyReactiveTokenAccount(TokenAccount):
    """Never proactive; reacts with ``int(utility * k)`` messages."""
    KIND = 1

    def __init__(self, k: int = 1) -> None:
        super().__init__()
        self.k = k

    def proactive(self) -> float:
        return 0

    def reactive(self, utility: int) -> int:
        return int(utility * self.k)


class SimpleTokenAccount(TokenAccount):
    """Proactive once the balance reaches ``C``; reacts with one message if any token."""
    KIND = 2

    def __init__(self, C: int = 1) -> None:
        super().__init__()
        assert C >= 1, "The capacity C must be strictly positive."
        self.capacity = C

    def proactive(self) -> float:
        return int(self.n_tokens >= self.capacity)

    def reactive(self, utility: int) -> int:
        return int(self.n_tokens > 0)


class GeneralizedTokenAccount(SimpleTokenAccount):
    """Reactive count ``floor((A-1+a)/A)`` (halved for useless messages)."""
    KIND = 3

    def __init__(self, C: int, A: int) -> None:
        super().__init__(C)
        assert A >= 1, "The reactivity A must be positive."
        assert A <= C, "The capacity C must be greater or equal than the reactivity A."
        self.reactivity = A

    def reactive(self, utility: int) -> int:
        num = self.reactivity + self.n_tokens - 1
        den = self.reactivity if utility > 0 else 2 * self.reactivity
        return int(num / den)


class RandomizedTokenAccount(GeneralizedTokenAccount):
    """Linear proactive ramp between ``A-1`` and ``C``; randomised-rounding reactive count."""
    KIND = 4

    def proactive(self) -> float:
        a, A, C = self.n_tokens, self.reactivity, self.capacity
        if a < A - 1:
            return 0
        if a <= C:
            return (a - A + 1) / (C - A + 1)
        return 1

    def reactive(self, utility: int) -> int:
        if utility <= 0:
            return 0
        r = self.n_tokens / self.reactivity
        whole = int(r)
        return whole + int(np.random.binomial(1, r - whole))
