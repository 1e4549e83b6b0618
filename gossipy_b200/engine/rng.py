"""Counter-based randomness shared by host and device code.

Every device-side random decision (mini-batch shuffles, sampled-merge coordinates) is a pure
function of ``(base_seed, stream, counter)`` so that a simulation is bit-reproducible however
nodes are placed on GPUs.  The same mixing function is implemented in
``csrc/common.cuh`` (``gb_mix64``); ``tests/test_rng.py`` checks that both agree.
"""
from __future__ import annotations

_MASK = (1 << 64) - 1
_base_seed = 0


def set_base_seed(seed: int) -> None:
    global _base_seed
    _base_seed = int(seed) & _MASK


def base_seed() -> int:
    return _base_seed


def mix64(x: int) -> int:
    """splitmix64 finaliser (bijective on 64-bit integers)."""
    x = (x + 0x9E3779B97F4A7C15) & _MASK
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & _MASK
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & _MASK
    return x ^ (x >> 31)


def derive(*parts: int) -> int:
    """Derive a 63-bit sub-seed from the base seed and a tuple of integers."""
    h = mix64(_base_seed)
    for p in parts:
        h = mix64(h ^ (int(p) & _MASK))
    return h & ((1 << 63) - 1)


def feistel_perm(i: int, n: int, key: int) -> int:
    """Index ``i`` of a keyed pseudo-random permutation of ``range(n)`` (cycle walking).

    Mirrors ``gb_perm`` in ``csrc/common.cuh``: a 4-round Feistel network on the smallest
    even-width power-of-two domain covering ``n``, re-applied until the value falls below n.
    """
    assert 0 <= i < n
    bits = max(2, (n - 1).bit_length())
    if bits & 1:
        bits += 1
    half = bits // 2
    hmask = (1 << half) - 1
    x = i
    while True:
        l, r = x >> half, x & hmask
        for rnd in range(4):
            f = mix64(key ^ (rnd << 56) ^ r) & hmask
            l, r = r, l ^ f
        x = (l << half) | r
        if x < n:
            return x
