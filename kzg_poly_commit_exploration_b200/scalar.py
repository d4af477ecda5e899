"""Host mirror of the reference's `scalar` module (src/scalar.rs): `Scalar` over Fr.

Same method names and meaning as the Rust type.  Values are kept canonical (32-byte LE, < r);
arithmetic goes through the C ABI's single-element helpers (kzg_fr_*), never through the oracle."""
from __future__ import annotations

import ctypes as C

from . import _native as N

R_MODULUS = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001  # src/scalar.rs:10


def _buf(b: bytes):
    return (C.c_uint8 * 32).from_buffer_copy(b)


class Scalar:
    __slots__ = ("_le",)

    def __init__(self, le_bytes: bytes = bytes(32)):
        o = N.out(32)
        N.lib().kzg_fr_reduce(_buf(le_bytes), o)   # inputs >= r are reduced (src/scalar.rs:54-61)
        self._le = bytes(o)

    # ---- constructors (src/scalar.rs:27-80) ---------------------------------------------
    @classmethod
    def from_i128(cls, a: int) -> "Scalar":
        """a > 0 -> a ; a <= 0 -> r - |a| (0 -> 0), like `impl From<i128>` (src/scalar.rs:27-48)."""
        if not -(1 << 127) <= a < (1 << 127):
            raise OverflowError("i128 out of range")
        mag = cls(abs(a).to_bytes(32, "little"))
        return mag if a > 0 else mag.neg()

    @classmethod
    def from_le_bytes(cls, b: bytes) -> "Scalar":
        if len(b) != 32:
            raise ValueError(f"Invalid length, expected 32, got {len(b)}")
        return cls(bytes(b))

    @classmethod
    def from_be_bytes(cls, b: bytes) -> "Scalar":
        if len(b) != 32:
            raise ValueError(f"Invalid length, expected 32, got {len(b)}")
        return cls(bytes(b)[::-1])

    # ---- export (src/scalar.rs:83-106) ---------------------------------------------------
    def to_le_bytes(self) -> bytes:
        return self._le

    def to_be_bytes(self) -> bytes:
        return self._le[::-1]

    # ---- arithmetic (src/scalar.rs:111-223) ----------------------------------------------
    def _bin(self, fn, other: "Scalar") -> "Scalar":
        o = N.out(32)
        fn(_buf(self._le), _buf(other._le), o)
        s = Scalar.__new__(Scalar)
        s._le = bytes(o)
        return s

    def mul(self, other: "Scalar") -> "Scalar":
        return self._bin(N.lib().kzg_fr_mul, other)

    def add(self, other: "Scalar") -> "Scalar":
        return self._bin(N.lib().kzg_fr_add, other)

    def sub(self, other: "Scalar") -> "Scalar":
        return self._bin(N.lib().kzg_fr_sub, other)

    def neg(self) -> "Scalar":
        o = N.out(32)
        N.lib().kzg_fr_neg(_buf(self._le), o)
        s = Scalar.__new__(Scalar)
        s._le = bytes(o)
        return s

    def pow(self, n: int) -> "Scalar":
        """self^n (src/scalar.rs:122-187; the value is canonical, whatever the multiplication chain)."""
        o = N.out(32)
        N.lib().kzg_fr_pow(_buf(self._le), n, o)
        s = Scalar.__new__(Scalar)
        s._le = bytes(o)
        return s

    def is_zero(self) -> bool:
        return self._le == bytes(32)

    def clone(self) -> "Scalar":
        return self

    def __eq__(self, other):
        return isinstance(other, Scalar) and self._le == other._le

    def __hash__(self):
        return hash(self._le)

    def __int__(self):
        return int.from_bytes(self._le, "little")

    def __str__(self):   # Display: base-10 (src/scalar.rs:277-341)
        return str(int(self))

    def __repr__(self):
        return f"Scalar({int(self)})"

    # serde: 32 LE bytes as a JSON array of integers (src/scalar.rs:226-275)
    def to_json(self):
        return list(self._le)

    @classmethod
    def from_json(cls, seq) -> "Scalar":
        if len(seq) != 32:
            raise ValueError(f"Invalid length, expected 32, got {len(seq)}")
        return cls(bytes(seq))
