"""Host mirror of the reference's `polynomial` module (src/polynomial.rs): `Polynomial`, `Evaluation`.

Same names, argument meaning and error behaviour as the Rust API; every bulk operation is one C-ABI
call into the CUDA engine (kzg_commit / kzg_evaluate / kzg_quotient / kzg_open)."""
from __future__ import annotations

from . import _native as N
from .curves import G1Point, G2Point, bilinear_maps_equal
from .scalar import Scalar
from .trusted_setup import TrustedSetup


class KzgError(Exception):
    """anyhow::Error of the reference; `.kind` names the C-ABI status."""

    def __init__(self, kind: str, msg: str):
        super().__init__(msg)
        self.kind = kind


def _wrap(e: N.KzgNativeError) -> KzgError:
    return KzgError(e.kind, str(e).split("] ", 1)[-1])


class Polynomial:
    def __init__(self, coefficients):
        self.coefficients = list(coefficients)

    # TryFrom<Vec<Scalar>> / TryFrom<Vec<i128>> / From<Scalar> (src/polynomial.rs:14-89)
    @classmethod
    def try_from(cls, value) -> "Polynomial":
        vals = [v if isinstance(v, Scalar) else Scalar.from_i128(v) for v in value]
        if len(vals) > 0xFFFFFFFF:
            raise KzgError("BAD_LENGTH", "Too many coefficients for polynomial, only 2**32 - 1 coefficients is supported.")
        last = 0
        for i, v in enumerate(vals):
            if not v.is_zero():
                last = i
        return cls(vals[: last + 1])

    @classmethod
    def from_scalar(cls, value: Scalar) -> "Polynomial":
        return cls([] if value.is_zero() else [value])

    def degree(self) -> int:                       # src/polynomial.rs:95-100
        return 0 if not self.coefficients else len(self.coefficients) - 1

    def _bytes(self) -> bytes:
        return b"".join(c.to_le_bytes() for c in self.coefficients)

    def evaluate(self, x: Scalar, ctx=None) -> "Evaluation":     # src/polynomial.rs:112-123
        from .trusted_setup import default_context

        ctx = ctx or default_context()
        y = ctx.evaluate(self._bytes(), x.to_le_bytes(), n=len(self.coefficients))
        return Evaluation(x, Scalar(y))

    def sub(self, other: "Polynomial") -> "Polynomial":          # src/polynomial.rs:128-145
        a, b = self.coefficients, other.coefficients
        if len(a) > len(b):
            co = list(a)
            for i, rhs in enumerate(b):
                co[i] = co[i].sub(rhs)
        else:
            co = [x.neg() for x in b]
            for i, lhs in enumerate(a):
                co[i] = lhs.add(co[i])
        return Polynomial.try_from(co)

    def divide_by_root(self, root: Scalar, ctx=None) -> "Polynomial":   # src/polynomial.rs:150-195
        from .trusted_setup import default_context

        ctx = ctx or default_context()
        if not self.coefficients:
            return Polynomial([])
        if len(self.coefficients) == 1:
            if self.coefficients[0].is_zero():
                return Polynomial([])
            raise KzgError("CONSTANT_POLY", "Unable to divide a constant polynomial")
        try:
            q = ctx.quotient(self._bytes(), root.to_le_bytes(), bytes(32), n=len(self.coefficients))
        except N.KzgNativeError as e:
            raise _wrap(e) from None
        return Polynomial.try_from([Scalar(q[i: i + 32]) for i in range(0, len(q), 32)])

    def commit(self, setup_artifacts: TrustedSetup) -> G1Point:  # src/polynomial.rs:200-215
        try:
            c = setup_artifacts.ctx.commit(setup_artifacts.dev, self._bytes(), n=len(self.coefficients))
        except N.KzgNativeError as e:
            raise _wrap(e) from None
        return G1Point._raw(c)

    def __str__(self):                              # Display (src/polynomial.rs:218-247)
        if not self.coefficients:
            return "0"

        def term(c, d):
            return f"{c}" + ("" if d == 0 else "x" if d == 1 else f"x^{d}")

        n = len(self.coefficients)
        out = term(self.coefficients[-1], n - 1)
        for i in range(n - 2, -1, -1):
            if not self.coefficients[i].is_zero():
                out += " + " + term(self.coefficients[i], i)
        return out

    def to_json(self):
        return {"coefficients": [c.to_json() for c in self.coefficients]}


class Evaluation:
    def __init__(self, point: Scalar, result: Scalar):
        self.point = point
        self.result = result

    def generate_proof(self, polynomial: Polynomial, setup_artifacts: TrustedSetup) -> G1Point:
        """src/polynomial.rs:260-269: commit((p - result) / (x - point)); the device recomputes p(point) and
        rejects a wrong `result` with the reference's divide_by_root errors."""
        try:
            _, proof = setup_artifacts.ctx.open(setup_artifacts.dev, polynomial._bytes(), self.point.to_le_bytes(),
                                                self.result.to_le_bytes(), n=len(polynomial.coefficients))
        except N.KzgNativeError as e:
            raise _wrap(e) from None
        return G1Point._raw(proof)

    def verify_proof(self, proof: G1Point, commitment: G1Point, setup_artifacts) -> bool:
        """src/polynomial.rs:276-294: e(proof, setup[1].g2 - [point]_2) == e(commitment - [result]_1, [1]_2).
        Indexing setup[1] fails on a setup shorter than 2 items, as the reference's does (:284)."""
        tau2 = setup_artifacts[1].g2
        if tau2 is None:
            raise KzgError("SETUP_TOO_SHORT", "setup item 1 carries no G2 point")
        lhs_g2 = tau2.sub(G2Point.from_scalar(self.point))
        rhs_g1 = commitment.sub(G1Point.from_scalar(self.result))
        return bilinear_maps_equal(proof, lhs_g2, rhs_g1, G2Point.from_i128(1))

    def to_json(self):
        return {"point": self.point.to_json(), "result": self.result.to_json()}
