"""Shared helpers for the parity tests (numpy-vectorised deterministic inputs)."""
from __future__ import annotations

import numpy as np

import kzg_oracle as O

_C1 = np.uint64(0xBF58476D1CE4E5B9)
_C2 = np.uint64(0x94D049BB133111EB)
_G = 0x9E3779B97F4A7C15


def _mix(z: np.ndarray) -> np.ndarray:
    z = (z ^ (z >> np.uint64(30))) * _C1
    z = (z ^ (z >> np.uint64(27))) * _C2
    return z ^ (z >> np.uint64(31))


def scalar_stream_bytes(seed: int, n: int, start: int = 0) -> bytes:
    """n x 32-byte LE scalars U(seed, start..start+n), vectorised; identical to
    kzg_oracle.scalar_stream."""
    with np.errstate(over="ignore"):
        k = np.arange(start, start + n, dtype=np.uint64)
        words = np.empty((n, 4), dtype=np.uint64)
        for j in range(4):
            ctr = (np.uint64(4) * k + np.uint64(j + 1)) * np.uint64(_G) + np.uint64(seed & 0xFFFFFFFFFFFFFFFF)
            words[:, j] = _mix(ctr)
        words[:, 3] &= np.uint64((1 << 63) - 1)
    raw = words.astype("<u8").tobytes()
    # subtract r where >= r (rare: 2^255 / r ~ 1.1 -> ~9.6% of values)
    out = bytearray(raw)
    r_words = [(O.R >> (64 * j)) & 0xFFFFFFFFFFFFFFFF for j in range(4)]
    ge = np.zeros(n, dtype=bool)
    eq = np.ones(n, dtype=bool)
    for j in (3, 2, 1, 0):
        w = words[:, j]
        ge |= eq & (w > np.uint64(r_words[j]))
        eq &= w == np.uint64(r_words[j])
    ge |= eq
    for i in np.nonzero(ge)[0]:
        v = int.from_bytes(raw[32 * i: 32 * i + 32], "little") - O.R
        out[32 * i: 32 * i + 32] = v.to_bytes(32, "little")
    return bytes(out)


def to_ints(b: bytes):
    return [int.from_bytes(b[i: i + 32], "little") for i in range(0, len(b), 32)]


def from_ints(vals) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def le(v: int) -> bytes:
    return int(v).to_bytes(32, "little")
