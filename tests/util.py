"""Shared helpers for the parity tests."""
from __future__ import annotations


def scalar_stream_bytes(seed: int, n: int, start: int = 0) -> bytes:
    """n x 32-byte LE scalars U(seed, start..start+n) from the product's generator, which is pinned
    against kzg_oracle.scalar_stream in tests/test_oracle.py."""
    from kzg_poly_commit_exploration_b200 import synthetic

    return synthetic.scalar_stream(seed, n, start).tobytes()


def to_ints(b: bytes):
    return [int.from_bytes(b[i: i + 32], "little") for i in range(0, len(b), 32)]


def from_ints(vals) -> bytes:
    return b"".join(int(v).to_bytes(32, "little") for v in vals)


def le(v: int) -> bytes:
    return int(v).to_bytes(32, "little")
