"""Host mirror of the reference's `curves` module (src/curves.rs): `G1Point`, `G2Point`, `bilinear_map`.

A point holds its canonical compressed encoding (48 / 96 bytes, what Serialize emits, :99-110, :260-271);
per-element group operations go through the C ABI's host helpers (kzg_g1_*, kzg_g2_*).  The pairing
(src/curves.rs:355-371) is two Miller loops whatever the degree: host code in the same library
(kzg_pairing_equal), exposed here in the only form the reference consumes — equality of two pairings."""
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


def _b96(b: bytes):
    return (C.c_uint8 * 96).from_buffer_copy(b)


class G2Point:
    """src/curves.rs:185-343."""
    __slots__ = ("_c",)

    def __init__(self, compressed: bytes):
        """Deserialize (src/curves.rs:273-343): length 96, valid flags, on the twist; no subgroup check."""
        if len(compressed) != 96:
            raise N.KzgNativeError(5, f"Invalid length, expected 96, got {len(compressed)}")
        if N.lib().kzg_g2_check(_b96(compressed)) != 0:
            raise N.KzgNativeError(4, "Got error while uncompressing")
        self._c = bytes(compressed)

    @classmethod
    def _raw(cls, c: bytes) -> "G2Point":
        p = cls.__new__(cls)
        p._c = c
        return p

    @classmethod
    def from_scalar(cls, a: Scalar) -> "G2Point":
        """[a]G2 (src/curves.rs:225-236)."""
        o = N.out(96)
        N.lib().kzg_g2_from_scalar((C.c_uint8 * 32).from_buffer_copy(a.to_le_bytes()), o)
        return cls._raw(bytes(o))

    @classmethod
    def from_i128(cls, a: int) -> "G2Point":
        """[|a|]G2, negated when a < 0 (src/curves.rs:203-220)."""
        p = cls.from_scalar(Scalar(abs(a).to_bytes(32, "little")))
        if a < 0:
            o = N.out(96)
            N.lib().kzg_g2_neg(_b96(p._c), o)
            return cls._raw(bytes(o))
        return p

    def sub(self, b: "G2Point") -> "G2Point":      # src/curves.rs:241-249
        o = N.out(96)
        st = N.lib().kzg_g2_sub(_b96(self._c), _b96(b._c), o)
        if st != 0:
            raise N.KzgNativeError(st, "Got error while uncompressing")
        return G2Point._raw(bytes(o))

    def compress(self) -> bytes:
        return self._c

    def to_json(self):
        return list(self._c)

    @classmethod
    def from_json(cls, seq) -> "G2Point":
        return cls(bytes(seq))

    def __eq__(self, other):
        return isinstance(other, G2Point) and self._c == other._c

    def __repr__(self):
        return f"G2Point({self._c.hex()})"


def bilinear_maps_equal(p1: G1Point, q1: G2Point, p2: G1Point, q2: G2Point) -> bool:
    """bilinear_map(p1, q1) == bilinear_map(p2, q2) (src/curves.rs:355-371, consumed only as an equality,
    src/polynomial.rs:293)."""
    eq = C.c_int32(0)
    st = N.lib().kzg_pairing_equal(_b48(p1._c), _b96(q1._c), _b48(p2._c), _b96(q2._c), C.byref(eq))
    if st != 0:
        raise N.KzgNativeError(st, "Got error while uncompressing")
    return bool(eq.value)
