"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/kzg_b200.h
declares, refuses to run bulk work without a CUDA device (no CPU fallback), and its host
single-element helpers — the host instantiation of the same field/curve code the kernels use —
agree with the oracle."""
import ctypes as C
import os
import random
import re

import pytest

import kzg_oracle as O
import util

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol(native):
    header = open(os.path.join(ROOT, "include", "kzg_b200.h")).read()
    declared = set(re.findall(r"\b(kzg_[a-z0-9_]+)\s*\(", header))
    declared -= {"kzg_status", "kzg_g1_partial", "kzg_g1_combine"}  # words in comments / the enum, not functions
    L = native.lib()
    missing = [n for n in sorted(declared) if not hasattr(L, n)]
    assert not missing, missing
    assert declared == set(native._native.SIGNATURES), declared ^ set(native._native.SIGNATURES)


def test_no_cpu_fallback_without_gpu(native):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(native.KzgNativeError) as ei:
        native.Context(0)
    assert ei.value.kind == "CUDA"


def test_status_strings_match_reference_messages(native):
    L = native.lib()
    assert b"degree is too high" in L.kzg_status_string(1)
    assert L.kzg_status_string(2) == b"Unable to divide a constant polynomial"
    assert b"constant terms do not add up" in L.kzg_status_string(3)


def _b(x, n=32):
    return (C.c_uint8 * n).from_buffer_copy(x)


def test_fr_host_helpers(native):
    L = native.lib()
    rnd = random.Random(5)
    o = (C.c_uint8 * 32)()
    for it in range(300):
        a, b = rnd.randrange(1 << 256), rnd.randrange(1 << 256)
        if it < 5:
            a = [0, 1, O.R - 1, O.R, (1 << 256) - 1][it]
        A, B = _b(util.le(a)), _b(util.le(b))
        L.kzg_fr_mul(A, B, o); assert int.from_bytes(bytes(o), "little") == a * b % O.R
        L.kzg_fr_add(A, B, o); assert int.from_bytes(bytes(o), "little") == (a + b) % O.R
        L.kzg_fr_sub(A, B, o); assert int.from_bytes(bytes(o), "little") == (a - b) % O.R
        L.kzg_fr_neg(A, o); assert int.from_bytes(bytes(o), "little") == (-a) % O.R
        L.kzg_fr_reduce(A, o); assert int.from_bytes(bytes(o), "little") == a % O.R
        n = rnd.randrange(1 << 24)
        L.kzg_fr_pow(A, n, o); assert int.from_bytes(bytes(o), "little") == pow(a, n, O.R)


def test_g1_host_helpers(native):
    L = native.lib()
    rnd = random.Random(6)
    o = (C.c_uint8 * 48)()
    inf = O.g1_compress(None)
    for it in range(12):
        k, k2 = rnd.randrange(1 << 256), rnd.randrange(O.R)
        if it < 5:
            k = [0, 1, 2, O.R - 1, O.R][it]
        L.kzg_g1_from_scalar(_b(util.le(k)), o)
        P = O.g1_mul(O.G1, k % O.R)
        assert bytes(o) == O.g1_compress(P)
        Q2 = O.g1_mul(O.G1, k2)
        cp, cq = _b(O.g1_compress(P), 48), _b(O.g1_compress(Q2), 48)
        assert L.kzg_g1_add(cp, cq, o) == 0 and bytes(o) == O.g1_compress(O.g1_add(P, Q2))
        assert L.kzg_g1_sub(cp, cq, o) == 0 and bytes(o) == O.g1_compress(O.g1_sub(P, Q2))
        assert L.kzg_g1_add(cp, cp, o) == 0 and bytes(o) == O.g1_compress(O.g1_add(P, P))      # doubling case
        assert L.kzg_g1_sub(cp, cp, o) == 0 and bytes(o) == inf                               # P - P
        assert L.kzg_g1_mul(cq, _b(util.le(k)), o) == 0 and bytes(o) == O.g1_compress(O.g1_mul(Q2, k % O.R))
    # Deserialize semantics (src/curves.rs:119-144): flags, x < q, on-curve; no subgroup check
    assert L.kzg_g1_check(_b(O.g1_compress(O.G1), 48)) == 0
    assert L.kzg_g1_check(_b(inf, 48)) == 0
    assert L.kzg_g1_check(_b(b"\x00" * 48, 48)) == 4
    assert L.kzg_g1_check(_b(bytes([0xC0]) + bytes(46) + b"\x01", 48)) == 4
    assert L.kzg_g1_check(_b(bytes([0x9F]) + b"\xff" * 47, 48)) == 4          # x >= q
    off = next(x for x in range(2, 50) if pow((x**3 + 4) % O.Q, (O.Q - 1) // 2, O.Q) != 1)
    enc = bytearray(off.to_bytes(48, "big")); enc[0] |= 0x80
    assert L.kzg_g1_check(_b(bytes(enc), 48)) == 4                              # not on the curve
    # a point of E(Fq) outside the r-torsion is ACCEPTED, as by blst_p1_uncompress
    x = next(x for x in range(1, 50) if pow((x**3 + 4) % O.Q, (O.Q - 1) // 2, O.Q) == 1)
    enc = bytearray(x.to_bytes(48, "big")); enc[0] |= 0x80
    assert L.kzg_g1_check(_b(bytes(enc), 48)) == 0
