"""Host mirror of the reference's `curves` module, G1 part (src/curves.rs:10-183): `G1Point`.

A G1Point holds the canonical 48-byte compressed encoding (what Serialize emits, :99-110);
per-element group operations go through the C ABI's host helpers (kzg_g1_*).  G2 and the pairing
(src/curves.rs:185-371) stay with blst in a Rust integration: verification is two
pairings whatever the degree and is not on the accelerated path."""
from __future__ import annotations

import ctypes as C

from . import _native as N
from .scalar import Scalar


def _b48(b: bytes):
    return (C.c_uint8 * 48).from_buffer_copy(b)


class G1Point:
    __slots__ = ("_c",)

    def __init__(self, compressed: bytes):
        """Deserialize (src/curves.rs:112-183): length 48, valid flags, on the curve; no subgroup check."""
        if len(compressed) != 48:
            raise N.KzgNativeError(5, f"Invalid length, expected 48, got {len(compressed)}")
        if N.lib().kzg_g1_check(_b48(compressed)) != 0:
            raise N.KzgNativeError(4, "Got error while uncompressing")
        self._c = bytes(compressed)

    @classmethod
    def _raw(cls, c: bytes) -> "G1Point":
        p = cls.__new__(cls)
        p._c = c
        return p

    @classmethod
    def from_scalar(cls, a: Scalar) -> "G1Point":
        """[a]G1 (src/curves.rs:50-61)."""
        o = N.out(48)
        N.lib().kzg_g1_from_scalar((C.c_uint8 * 32).from_buffer_copy(a.to_le_bytes()), o)
        return cls._raw(bytes(o))

    @classmethod
    def from_i128(cls, a: int) -> "G1Point":
        """[|a|]G1, negated when a < 0 (src/curves.rs:28-45)."""
        p = cls.from_scalar(Scalar(abs(a).to_bytes(32, "little")))
        return cls.from_i128(0).sub(p) if a < 0 else p

    def _bin(self, fn, other: "G1Point") -> "G1Point":
        o = N.out(48)
        st = fn(_b48(self._c), _b48(other._c), o)
        if st != 0:
            raise N.KzgNativeError(st, "Got error while uncompressing")
        return G1Point._raw(bytes(o))

    def add(self, b: "G1Point") -> "G1Point":      # src/curves.rs:79-85
        return self._bin(N.lib().kzg_g1_add, b)

    def sub(self, b: "G1Point") -> "G1Point":      # src/curves.rs:66-74
        return self._bin(N.lib().kzg_g1_sub, b)

    def mult(self, a: Scalar) -> "G1Point":        # src/curves.rs:90-96
        o = N.out(48)
        st = N.lib().kzg_g1_mul(_b48(self._c), (C.c_uint8 * 32).from_buffer_copy(a.to_le_bytes()), o)
        if st != 0:
            raise N.KzgNativeError(st, "Got error while uncompressing")
        return G1Point._raw(bytes(o))

    def compress(self) -> bytes:
        return self._c

    def to_json(self):                              # serialize_bytes -> JSON array of ints (src/curves.rs:108)
        return list(self._c)

    @classmethod
    def from_json(cls, seq) -> "G1Point":
        return cls(bytes(seq))

    def __eq__(self, other):
        return isinstance(other, G1Point) and self._c == other._c

    def __repr__(self):
        return f"G1Point({self._c.hex()})"
