"""The reference's own unit tests, restated against the host mirror of its public API
(kzg_poly_commit_exploration_b200.{scalar,curves,polynomial,trusted_setup} — same names as the Rust
modules).  CPU-only ones exercise the host helpers of the C ABI; the KZG flow needs the GPU."""
import random

import pytest

import kzg_oracle as O
from kzg_poly_commit_exploration_b200.curves import G1Point, G2Point, bilinear_maps_equal
from kzg_poly_commit_exploration_b200.polynomial import Evaluation, KzgError, Polynomial
from kzg_poly_commit_exploration_b200.scalar import Scalar


# ---- src/scalar.rs:350-414 -------------------------------------------------------------------------
def test_i128_to_scalar_using_le():
    rnd = random.Random(1)
    for _ in range(100):
        a = rnd.randrange(-(1 << 127), 1 << 127)
        expected = a if a > 0 else (O.R - (-a)) % O.R
        assert Scalar.from_i128(a).to_le_bytes() == expected.to_bytes(32, "little")


def test_u128_to_scalar_using_le_and_be():
    rnd = random.Random(2)
    for _ in range(100):
        a = rnd.randrange(1 << 128)
        assert Scalar.from_le_bytes(a.to_bytes(32, "little")).to_le_bytes() == a.to_bytes(32, "little")
        assert Scalar.from_be_bytes(a.to_bytes(32, "big")).to_be_bytes() == a.to_bytes(32, "big")


def test_display_scalar_and_pow():
    rnd = random.Random(3)
    for _ in range(50):
        a = rnd.randrange(1 << 256) % O.R
        assert str(Scalar.from_le_bytes(a.to_bytes(32, "little"))) == str(a)
        base, e = rnd.randrange(10**6), rnd.randrange(10)
        assert int(Scalar.from_i128(base).pow(e)) == base**e % O.R
    assert int(Scalar.from_i128(5).pow(0)) == 1


# ---- src/curves.rs:375-451 ---------------------------------------------------------------------------
def test_point_addition_and_scalar_multiplication():
    g = G1Point.from_i128(1)
    assert g.add(g) == G1Point.from_i128(2) == g.mult(Scalar.from_i128(2))
    assert G1Point.from_i128(-1).compress() == O.g1_compress(O.g1_neg(O.G1))
    assert g.sub(g).compress() == O.g1_compress(None)


def test_compression_and_serialization_roundtrip():
    two = G1Point.from_i128(2)
    assert G1Point.from_json(two.to_json()) == two
    with pytest.raises(Exception):
        G1Point(bytes(47))
    with pytest.raises(Exception):
        G1Point(bytes(48))


def test_g2_point_mirror():
    g = G2Point.from_i128(1)
    assert g.compress() == O.g2_compress(O.G2)
    assert G2Point.from_i128(-3).compress() == O.g2_compress(O.g2_neg(O.g2_mul(O.G2, 3)))
    assert G2Point.from_scalar(Scalar.from_i128(5)).sub(G2Point.from_i128(2)) == G2Point.from_i128(3)
    assert G2Point.from_json(g.to_json()) == g
    with pytest.raises(Exception):
        G2Point(bytes(96))
    with pytest.raises(Exception):
        G2Point(bytes(95))
    a, b = Scalar.from_i128(77), Scalar.from_i128(1001)
    assert bilinear_maps_equal(G1Point.from_scalar(a), G2Point.from_scalar(b),
                               G1Point.from_scalar(a.mul(b)), G2Point.from_i128(1))


# ---- src/polynomial.rs:301-321 -----------------------------------------------------------------------
def test_create_polynomial_with_tailing_zeros():
    p = Polynomial.try_from([1, 2, 3, 0, 0])
    assert len(p.coefficients) == 3 and p.degree() == 2
    z = Polynomial.try_from([0, 0, 0])
    assert len(z.coefficients) == 1 and z.degree() == 0
    assert Polynomial.from_scalar(Scalar.from_i128(0)).coefficients == []
    assert str(Polynomial.try_from([2, 1, 3])) == "3x^2 + 1x + 2"


# ---- src/lib.rs:16-94: commit -> evaluate -> prove -> verify ---------------------------------------------
def _kate_roundtrip(ctx, secret, coeffs_i128, point_i128):
    from kzg_poly_commit_exploration_b200.trusted_setup import SetupArtifactsGenerator

    poly = Polynomial.try_from(coeffs_i128)
    setup = SetupArtifactsGenerator(secret, ctx).take(poly.degree() + 1)
    x = Scalar.from_i128(point_i128)
    commitment = poly.commit(setup)
    evaluation = poly.evaluate(x, ctx)
    proof = evaluation.generate_proof(poly, setup)
    # Evaluation::verify_proof through the library's own pairing (src/lib.rs:26-32) ...
    assert evaluation.verify_proof(proof, commitment, setup)
    assert not Evaluation(x, evaluation.result.add(Scalar.from_i128(1))).verify_proof(proof, commitment, setup)
    # ... and, independently, through the oracle's pairing: e(pi,[s-z]2) == e(C-[y]1,G2)
    so = O.setup_generate(secret, max(2, poly.degree() + 1))
    ev = O.Evaluation(int(x), int(evaluation.result))
    assert ev.verify_proof(O.g1_uncompress(proof.compress()), O.g1_uncompress(commitment.compress()), so)
    return poly, setup, evaluation, proof


@pytest.mark.gpu
def test_kate_proof_for_polynomial_degree_one_and_two(ctx):
    rnd = random.Random(4)
    secret = rnd.randbytes(32)
    for degree in (1, 2):
        for _ in range(2):
            co = [rnd.randrange(-(1 << 127), 1 << 127) for _ in range(degree + 1)]
            _kate_roundtrip(ctx, secret, co, rnd.randrange(-(1 << 127), 1 << 127))


@pytest.mark.gpu
def test_kate_proof_over_multiple_degree_with_fixed_input(ctx):
    rnd = random.Random(5)
    point = rnd.randrange(-(1 << 127), 1 << 127)
    for degree in (37, 1999):
        secret = rnd.randbytes(32)
        co = [rnd.randrange(-(1 << 127), 1 << 127) for _ in range(degree + 1)]
        poly, setup, evaluation, proof = _kate_roundtrip(ctx, secret, co, point)
        # a tampered evaluation is rejected the way divide_by_root rejects it (src/polynomial.rs:188-192)
        bad = Evaluation(evaluation.point, evaluation.result.add(Scalar.from_i128(1)))
        with pytest.raises(KzgError) as ei:
            bad.generate_proof(poly, setup)
        assert ei.value.kind == "REMAINDER_NONZERO"
        # setup too short for the polynomial (src/polynomial.rs:201-205)
        with pytest.raises(KzgError) as ei:
            poly.commit(setup[: degree])
        assert ei.value.kind == "DEGREE_TOO_HIGH"
        # sub / divide_by_root mirror
        q = poly.sub(Polynomial.from_scalar(evaluation.result)).divide_by_root(evaluation.point, ctx)
        assert q.degree() == degree - 1 and q.commit(setup) == proof
