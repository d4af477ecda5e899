"""Deterministic synthetic inputs: the counter-based scalar stream U(seed, k) —
four SplitMix64 words masked to 255 bits, minus r if >= r — vectorised with numpy.  The same
stream is restated in oracle/kzg_oracle.py (scalar_stream) for the checker."""
from __future__ import annotations

import numpy as np

R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BENCH_SECRET_BE = bytes(range(32))  # benches/polynomial_commitment.rs:17-19
_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)
_G = np.uint64(0x9E3779B97F4A7C15)


def _mix(z):
    z = (z ^ (z >> np.uint64(30))) * _C1
    z = (z ^ (z >> np.uint64(27))) * _C2
    return z ^ (z >> np.uint64(31))


def scalar_stream(seed: int, n: int, start: int = 0) -> np.ndarray:
    """(n, 32) uint8 array: scalars U(seed, start .. start+n) as 32-byte little-endian values < r."""
    with np.errstate(over="ignore"):
        k = np.arange(start, start + n, dtype=np.uint64)
        words = np.empty((n, 4), dtype=np.uint64)
        for j in range(4):
            words[:, j] = _mix((np.uint64(4) * k + np.uint64(j + 1)) * _G + np.uint64(seed & 0xFFFFFFFFFFFFFFFF))
        words[:, 3] &= np.uint64((1 << 63) - 1)
        r_words = [np.uint64((R >> (64 * j)) & 0xFFFFFFFFFFFFFFFF) for j in range(4)]
        ge = np.zeros(n, dtype=bool)
        eq = np.ones(n, dtype=bool)
        for j in (3, 2, 1, 0):
            ge |= eq & (words[:, j] > r_words[j])
            eq &= words[:, j] == r_words[j]
        ge |= eq
        idx = np.nonzero(ge)[0]
        if len(idx):  # multi-word subtract of r on the few rows that need it
            borrow = np.zeros(len(idx), dtype=np.uint64)
            for j in range(4):
                a = words[idx, j]
                t = a - r_words[j]
                b1 = (a < r_words[j]).astype(np.uint64)
                t2 = t - borrow
                b2 = (t < borrow).astype(np.uint64)
                words[idx, j] = t2
                borrow = b1 | b2
    return np.ascontiguousarray(words.astype("<u8")).view(np.uint8).reshape(n, 32)
