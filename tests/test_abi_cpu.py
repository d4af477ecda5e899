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


# ---- G2 + pairing host helpers (src/curves.rs:185-371, src/polynomial.rs:276-294) ---------------------
def test_g2_host_helpers_match_oracle(native):
    L = native.lib()
    rnd = random.Random(9)
    o = (C.c_uint8 * 96)()
    # published generator encoding + small multiples + random scalars
    L.kzg_g2_from_scalar(_b(util.le(1)), o)
    assert bytes(o) == O.g2_compress(O.G2)
    assert bytes(o).hex().startswith("93e02b6052719f607dacd3a088274f65")
    ks = [0, 2, 3, O.R - 1, O.R, O.R + 5] + [rnd.randrange(1 << 256) for _ in range(4)]
    pts = []
    for k in ks:
        L.kzg_g2_from_scalar(_b(util.le(k)), o)
        exp = O.g2_compress(O.g2_from_scalar(k))
        assert bytes(o) == exp, k
        assert L.kzg_g2_check(_b(exp, 96)) == 0
        pts.append((k % O.R, exp))
    for (ka, a), (kb, b) in zip(pts, pts[1:] + pts[:1]):
        assert L.kzg_g2_add(_b(a, 96), _b(b, 96), o) == 0
        assert bytes(o) == O.g2_compress(O.g2_from_scalar((ka + kb) % O.R))
        assert L.kzg_g2_sub(_b(a, 96), _b(b, 96), o) == 0
        assert bytes(o) == O.g2_compress(O.g2_from_scalar((ka - kb) % O.R))
        assert L.kzg_g2_add(_b(a, 96), _b(a, 96), o) == 0           # doubling through the complete addition
        assert bytes(o) == O.g2_compress(O.g2_from_scalar(2 * ka % O.R))
        assert L.kzg_g2_sub(_b(a, 96), _b(a, 96), o) == 0           # P - P = infinity
        assert bytes(o) == bytes([0xC0]) + bytes(95)
        assert L.kzg_g2_neg(_b(a, 96), o) == 0
        assert bytes(o) == O.g2_compress(O.g2_from_scalar((-ka) % O.R))


def test_g2_uncompress_rejects_malformed(native):
    L = native.lib()
    good = bytearray(O.g2_compress(O.g2_from_scalar(7)))
    assert L.kzg_g2_check(_b(bytes(good), 96)) == 0
    bad = bytearray(good); bad[0] &= 0x7F                       # compression flag cleared
    assert L.kzg_g2_check(_b(bytes(bad), 96)) == 4
    bad = bytearray([0xC0]) + bytearray(95); bad[50] = 1        # infinity with payload
    assert L.kzg_g2_check(_b(bytes(bad), 96)) == 4
    bad = bytearray(b"\x9f" + b"\xff" * 95)                     # x.c1 >= q
    assert L.kzg_g2_check(_b(bytes(bad), 96)) == 4
    # an x with no point on the twist: walk x.c0 until the oracle's curve equation has no root
    found = False
    for d in range(1, 40):
        cand = bytearray(good)
        cand[95] = (cand[95] + d) & 0xFF
        st = L.kzg_g2_check(_b(bytes(cand), 96))
        if st == 4:
            found = True
            break
    assert found
    # both signs of a valid x round-trip to themselves
    o = (C.c_uint8 * 96)()
    flip = bytearray(good); flip[0] ^= 0x20
    assert L.kzg_g2_neg(_b(bytes(good), 96), o) == 0 and bytes(o) == bytes(flip)


def test_pairing_and_verify_proof(native):
    L = native.lib()
    eq = C.c_int32(-1)
    g1 = lambda k: _b(O.g1_compress(O.g1_from_scalar(k)), 48)
    g2 = lambda k: _b(O.g2_compress(O.g2_from_scalar(k)), 96)
    # bilinearity: e([a]G1, [b]G2) == e([ab]G1, G2) ; and a mismatch
    a, b = 0x1234567, 0xABCDEF0123
    assert L.kzg_pairing_equal(g1(a), g2(b), g1(a * b % O.R), g2(1), C.byref(eq)) == 0 and eq.value == 1
    assert L.kzg_pairing_equal(g1(a), g2(b), g1(a * b + 1), g2(1), C.byref(eq)) == 0 and eq.value == 0
    assert L.kzg_pairing_equal(g1(a), g2(b), g1(b), g2(a), C.byref(eq)) == 0 and eq.value == 1
    assert L.kzg_pairing_equal(g1(0), g2(b), g1(5), g2(0), C.byref(eq)) == 0 and eq.value == 1     # both sides 1
    # the reference's acceptance test (src/lib.rs:16-33) on the degree-100 bench inputs, values from the oracle
    s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
    co = O.bench_polynomial(100)
    z = O.bench_input_point(100)
    y = O.horner(co, z)
    commitment = O.g1_compress(O.commit_known_secret(co, s))
    proof = O.g1_compress(O.proof_known_secret(co, s, z))
    tau2 = O.g2_compress(O.g2_from_scalar(s))
    ok = C.c_int32(-1)
    args = lambda pr, cm, zz, yy: (_b(pr, 48), _b(cm, 48), _b(util.le(zz)), _b(util.le(yy)), _b(tau2, 96), C.byref(ok))
    assert L.kzg_verify_proof(*args(proof, commitment, z, y)) == 0 and ok.value == 1
    assert L.kzg_verify_proof(*args(proof, commitment, z, (y + 1) % O.R)) == 0 and ok.value == 0    # tampered evaluation
    assert L.kzg_verify_proof(*args(commitment, commitment, z, y)) == 0 and ok.value == 0          # tampered proof
    assert L.kzg_verify_proof(*args(bytes(48), commitment, z, y)) == 4                              # malformed point
    # agreement with the oracle's own pairing on one instance
    assert O.Evaluation(z, y).verify_proof(O.g1_uncompress(proof), O.g1_uncompress(commitment),
                                           O.setup_generate(O.BENCH_SECRET_BE, 2))
