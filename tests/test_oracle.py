"""Pins the oracle: published BLS12-381 encodings, the reference's own Fr encoding tests restated
(src/scalar.rs:350-414), its polynomial tests (src/polynomial.rs:301-321), the derived
known-answer vectors from the reference's bench inputs (tests/golden/kat_bench_inputs.json)
and the reference's acceptance test — commit -> evaluate -> prove -> pairing check (src/lib.rs:16-94)."""
import json
import os
import random

import pytest

import kzg_oracle as O

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "kat_bench_inputs.json")


def test_curve_constants_and_published_encodings():
    assert O.g1_is_on_curve(O.G1) and O.g2_is_on_curve(O.G2)
    assert O.g1_mul(O.G1, O.R) is None and O.g2_mul(O.G2, O.R) is None
    x = -O.BLS_X
    assert O.R == x**4 - x**2 + 1 and O.Q == (x - 1) ** 2 * (x**4 - x**2 + 1) // 3 + x
    assert O.g1_compress(O.G1).hex() == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    assert O.g1_compress(O.g1_mul(O.G1, 2)).hex() == "a572cbea904d67468808c8eb50a9450c9721db309128012543902d0ac358a62ae28f75bb8f1c7c42c39a8c5529bf0f4e"
    assert O.g1_compress(O.g1_neg(O.G1)).hex() == "b7f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"
    assert O.g1_compress(None) == bytes([0xC0]) + bytes(47)
    assert O.g2_compress(O.G2).hex() == (
        "93e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"
        "024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8")


def test_point_addition_and_compression_roundtrip():
    # src/curves.rs:375-451: G+G == [2]G; compress -> uncompress round trip
    two = O.g1_add(O.G1, O.G1)
    assert two == O.g1_mul(O.G1, 2)
    for k in [1, 2, 3, 12345, O.R - 1]:
        P = O.g1_mul(O.G1, k)
        assert O.g1_uncompress(O.g1_compress(P)) == P
    assert O.g1_uncompress(O.g1_compress(None)) is None
    with pytest.raises(O.KzgError):
        O.g1_uncompress(b"\x00" * 47)
    with pytest.raises(O.KzgError):
        O.g1_uncompress(b"\x00" * 48)


def test_i128_and_byte_encodings():
    # src/scalar.rs:350-389
    rnd = random.Random(1)
    for _ in range(200):
        a = rnd.randrange(-(1 << 127), 1 << 127)
        le = O.fr_to_le_bytes(O.fr_from_i128(a))
        assert int.from_bytes(le, "little") == (a if a > 0 else (O.R - (-a)) % O.R)
        u = rnd.randrange(1 << 128)
        assert O.fr_to_le_bytes(O.fr_from_le_bytes(u.to_bytes(32, "little"))) == u.to_bytes(32, "little")
        assert O.fr_to_be_bytes(O.fr_from_be_bytes(u.to_bytes(32, "big"))) == u.to_bytes(32, "big")
    assert O.fr_from_i128(0) == 0 and O.fr_from_i128(-(1 << 127)) == O.R - (1 << 127)
    assert O.fr_from_le_bytes((O.R + 5).to_bytes(32, "little")) == 5


def test_pow_restatement_and_mul_counts():
    # src/scalar.rs:403-414 plus the op counts SURVEY §3.3 quotes for evaluate
    rnd = random.Random(2)
    for _ in range(100):
        b, e = rnd.randrange(O.R), rnd.randrange(0, 5000)
        assert O.fr_pow_ref(b, e) == pow(b, e, O.R)
    cnt = [0]
    for d in range(1024):
        O.fr_pow_ref(3, d, cnt)
    assert 12.5 < cnt[0] / 1024 < 13.6   # + the c_i * x^i product = the survey's 13.5 Fr-mul/coefficient


def test_polynomial_trimming_and_zero_forms():
    # src/polynomial.rs:301-321 and SURVEY §8c edge cases 1-2
    p = O.Polynomial.try_from_i128([1, 2, 3, 0, 0])
    assert len(p.coefficients) == 3 and p.degree() == 2
    z = O.Polynomial.try_from_i128([0, 0, 0])
    assert z.coefficients == [0] and z.degree() == 0
    assert O.Polynomial.from_scalar(0).coefficients == [] and O.Polynomial.try_from_scalars([]).coefficients == []
    st = O.setup_generate(O.BENCH_SECRET_BE, 1)
    assert O.g1_compress(z.commit(st)) == O.g1_compress(None)
    with pytest.raises(O.KzgError) as ei:
        O.Polynomial.try_from_i128([1, 2]).commit(st)
    assert ei.value.kind == "DEGREE_TOO_HIGH"
    with pytest.raises(O.KzgError) as ei:
        O.Polynomial.try_from_i128([5]).divide_by_root(3)
    assert ei.value.kind == "CONSTANT_POLY"
    with pytest.raises(O.KzgError) as ei:
        O.Polynomial.try_from_i128([5, 1]).divide_by_root(3)
    assert ei.value.kind == "REMAINDER_NONZERO"
    assert O.Polynomial.try_from_i128([-15, 1, 0]).divide_by_root(15).coefficients == [1]


def test_known_answer_vectors_from_bench_inputs():
    with open(GOLDEN) as f:
        kat = json.load(f)
    s = O.fr_from_be_bytes(O.BENCH_SECRET_BE)
    st = O.setup_generate(O.BENCH_SECRET_BE, 3)
    assert [O.g1_compress(a.g1).hex() for a in st] == kat["setup_g1"][:3]
    assert O.g2_compress(st[1].g2).hex() == kat["setup1_g2"]
    for deg, v in kat["degrees"].items():
        deg = int(deg)
        co, z = O.bench_polynomial(deg), O.bench_input_point(deg)
        if deg <= 100:   # the reference's naive loop, restated
            stp = O.setup_generate(O.BENCH_SECRET_BE, deg + 1)
            p = O.Polynomial.try_from_scalars(co)
            ev = p.evaluate(z)
            assert O.g1_compress(p.commit(stp)).hex() == v["commit"]
            assert O.fr_to_le_bytes(ev.result).hex() == v["y_le"]
            assert O.fr_to_le_bytes(p.evaluate_ref_pow(z).result).hex() == v["y_le"]
            assert O.g1_compress(ev.generate_proof(p, stp)).hex() == v["proof"]
        assert O.g1_compress(O.commit_known_secret(co, s)).hex() == v["commit"]
        assert O.fr_to_le_bytes(O.horner(co, z)).hex() == v["y_le"]
        assert O.g1_compress(O.proof_known_secret(co, s, z)).hex() == v["proof"]


def test_kate_proof_pairing_acceptance():
    """src/lib.rs:51-94: random polynomials / points, verify == true; tampering -> false."""
    rnd = random.Random(7)
    secret = rnd.randbytes(32)
    for degree in (1, 2, 9):
        st = O.setup_generate(secret, degree + 1)
        p = O.Polynomial.try_from_i128([rnd.randrange(-(1 << 127), 1 << 127) for _ in range(degree + 1)])
        z = O.fr_from_i128(rnd.randrange(-(1 << 127), 1 << 127))
        c = p.commit(st)
        ev = p.evaluate(z)
        pr = ev.generate_proof(p, st)
        assert ev.verify_proof(pr, c, st)
        if degree == 2:
            assert not O.Evaluation(z, ev.result + 1).verify_proof(pr, c, st)
            assert not ev.verify_proof(O.g1_add(pr, O.G1), c, st)


def test_scalar_stream_matches_product_generator():
    from kzg_poly_commit_exploration_b200 import synthetic

    a = synthetic.scalar_stream(0xC3, 3000, 17)
    for i in range(3000):
        assert int.from_bytes(a[i].tobytes(), "little") == O.scalar_stream(0xC3, 17 + i)


def test_bench_commitment_golden_is_what_bench_asserts():
    """bench.py compares every timed MSM result with tests/golden/bench_commitment.json's value (embedded as a constant)."""
    import json
    import re
    with open(os.path.join(os.path.dirname(__file__), "golden", "bench_commitment.json")) as f:
        g = json.load(f)
    src = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "bench.py")).read()
    m = re.search(r'GOLDEN_C3_COMMITMENT = "([0-9a-f]{96})"', src)
    assert m and m.group(1) == g["commit"]
