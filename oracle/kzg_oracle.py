"""CPU oracle for the KZG hot path — TEST INFRASTRUCTURE ONLY.

This is a pure-Python (bigint) restatement of what the reference crate
VGLoic/kzg-poly-commit-exploration computes on its hot path.  It is the checker
for the CUDA engine; nothing under `kzg_poly_commit_exploration_b200/` may
import it.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` leg use it.

Parity status: the reference delegates all arithmetic to the third-party crate
`blst` 0.3.15 (Cargo.toml:10, Cargo.lock:89-92), whose C/assembly source is NOT
under /root/reference and cannot be built here (no cargo/rustc).  The
reference's own tests pin no commitment / proof / setup bytes
(src/lib.rs:16-94 only checks pairing self-consistency; src/scalar.rs:350-414
pins Fr byte encodings against num_bigint).  The oracle is therefore anchored
on
  (1) the published BLS12-381 / ZCash serialisation spec (well-known encodings
      of G1, 2*G1, -G1, G2 and infinity — checked in tests/test_oracle.py),
  (2) the reference's Fr encoding tests, restated in tests/test_oracle.py,
  (3) the reference's own acceptance test: commit -> evaluate -> prove ->
      pairing check == true (src/lib.rs:16-33), run with THIS oracle's pairing,
  (4) known-answer vectors derived from the reference's deterministic bench
      inputs (benches/polynomial_commitment.rs:10-23,
      benches/evaluation_proof.rs:10-27) by the survey's independent script
       and re-derived here (tests/golden/kat_bench_inputs.json).
Every output at the boundary is a canonical encoding (32-byte LE Fr < r,
48-byte compressed G1) of a mathematically unique value, so any correct
implementation is bit-identical to blst.  G1 outputs are "parity pinned by
spec + pairing", not by bytes shipped in the reference: by the letter of the
task's definition the G1 / G2 byte strings are "parity unpinned" by
reference-shipped vectors (none exist, and the crate cannot run here); the Fr
side is pinned by the reference's own encoding tests.

Each function cites the reference file:line it follows.
"""
from __future__ import annotations

# --------------------------------------------------------------------------
# Curve constants (public BLS12-381 parameters; reference states r at
# src/scalar.rs:10 and q in README.md:30)
# --------------------------------------------------------------------------
Q = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
BLS_X = 0xD201000000010000  # |x|, x is negative
G1_X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1_Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1
G2_X = (0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
        0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E)
G2_Y = (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
        0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE)
HALF_Q = (Q - 1) // 2


class KzgError(Exception):
    """Mirror of the reference's anyhow::Error; `.kind` matches the C-ABI status names."""

    def __init__(self, kind: str, msg: str):
        super().__init__(msg)
        self.kind = kind


# --------------------------------------------------------------------------
# Fr scalars — src/scalar.rs
# --------------------------------------------------------------------------
def fr_from_i128(a: int) -> int:
    """Scalar::from(i128): a>0 -> a ; a<=0 -> r-|a| with 0 -> 0 (src/scalar.rs:27-48)."""
    assert -(1 << 127) <= a < (1 << 127)
    if a > 0:
        return a % R
    return (R - (-a)) % R


def fr_from_le_bytes(b: bytes) -> int:
    """Scalar::from_le_bytes: 32 LE bytes, values >= r are reduced (src/scalar.rs:54-61)."""
    assert len(b) == 32
    return int.from_bytes(b, "little") % R


def fr_from_be_bytes(b: bytes) -> int:
    """Scalar::from_be_bytes (src/scalar.rs:66-73)."""
    assert len(b) == 32
    return int.from_bytes(b, "big") % R


def fr_to_le_bytes(a: int) -> bytes:
    """Scalar::to_le_bytes — canonical residue, 32 LE bytes (src/scalar.rs:83-93)."""
    return (a % R).to_bytes(32, "little")


def fr_to_be_bytes(a: int) -> bytes:
    """Scalar::to_be_bytes (src/scalar.rs:96-106)."""
    return (a % R).to_bytes(32, "big")


def fr_pow_ref(x: int, n: int, counter: list | None = None) -> int:
    """Scalar::pow — the reference's bespoke square-and-multiply, restated step
    by step (src/scalar.rs:122-187).  `counter[0]` accumulates Fr multiplications."""
    def mul(a, b):
        if counter is not None:
            counter[0] += 1
        return a * b % R

    if n == 0:
        return 1
    if n == 1:
        return x % R
    target = n // 2
    powered = []
    cur = 1
    while cur * 2 <= target:
        last = powered[-1] if powered else x
        powered.append(mul(last, last))
        cur *= 2
    if target % 2 == 0:
        acc, tracker = 1, 0
    else:
        acc, tracker = x % R, 1
    avail = 2 ** len(powered)
    while tracker != target:
        if tracker + avail <= target:
            tracker += avail
            acc = mul(acc, powered[-1])
        powered.pop()
        avail //= 2
    if target * 2 == n:
        return mul(acc, acc)
    return mul(mul(acc, acc), x)


# --------------------------------------------------------------------------
# G1 — affine tuples (x, y) or None for infinity; Jacobian internally.
# src/curves.rs:10-183
# --------------------------------------------------------------------------
G1 = (G1_X, G1_Y)


def g1_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - 4) % Q == 0


def g1_neg(P):
    if P is None:
        return None
    return (P[0], (-P[1]) % Q)


def _jac_double(P):
    X, Y, Z = P
    if Z == 0 or Y == 0:
        return (1, 1, 0)
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jac_add(P, Qp):
    """Complete Jacobian addition — same exceptional cases blst_p1_add_or_double
    absorbs (src/curves.rs:82): infinity operands, P == Q, P == -Q."""
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Qp
    if Z1 == 0:
        return Qp
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % Q
    Rr = (S2 - S1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (Rr * Rr - HHH - 2 * V) % Q
    Y3 = (Rr * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 * H % Q
    return (X3, Y3, Z3)


def _to_jac(P):
    return (1, 1, 0) if P is None else (P[0], P[1], 1)


def _from_jac(P):
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, Q - 2, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def g1_add(P, Qp):
    """G1Point::add (src/curves.rs:79-85)."""
    return _from_jac(_jac_add(_to_jac(P), _to_jac(Qp)))


def g1_sub(P, Qp):
    """G1Point::sub (src/curves.rs:66-74)."""
    return g1_add(P, g1_neg(Qp))


def _jac_mul(P, k: int):
    acc = (1, 1, 0)
    if k == 0 or P[2] == 0:
        return acc
    for bit in bin(k)[2:]:
        acc = _jac_double(acc)
        if bit == "1":
            acc = _jac_add(acc, P)
    return acc


def g1_mul(P, k: int):
    """G1Point::mult — 256-bit scalar multiplication by the canonical integer
    of the scalar (src/curves.rs:90-96).  NOTE: k is used as an integer, not
    reduced mod r here (blst_p1_mult consumes the 32 LE bytes as given; the
    bytes come from Scalar::to_le_bytes and so are already < r)."""
    return _from_jac(_jac_mul(_to_jac(P), k))


def g1_from_scalar(a: int):
    """G1Point::from_scalar (src/curves.rs:50-61)."""
    return g1_mul(G1, a % R)


def g1_from_i128(a: int):
    """G1Point::from_i128: [|a|]G, negated when a<0 (src/curves.rs:28-45)."""
    P = g1_mul(G1, abs(a))
    return g1_neg(P) if a < 0 else P


def g1_compress(P) -> bytes:
    """blst_p1_compress as used by Serialize for G1Point (src/curves.rs:99-110):
    ZCash format, 48 bytes big-endian x, bit7 compressed, bit6 infinity, bit5 y > (q-1)/2."""
    if P is None:
        return bytes([0xC0]) + bytes(47)
    x, y = P
    out = bytearray(x.to_bytes(48, "big"))
    out[0] |= 0x80
    if y > HALF_Q:
        out[0] |= 0x20
    return bytes(out)


def g1_uncompress(b: bytes):
    """bytes_to_blst_p1 (src/curves.rs:119-144): length check, blst_p1_uncompress
    (on-curve, NO subgroup check).  Raises KzgError('BAD_LENGTH'|'BAD_POINT')."""
    if len(b) != 48:
        raise KzgError("BAD_LENGTH", f"Invalid length, expected 48, got {len(b)}")
    if not (b[0] & 0x80):
        raise KzgError("BAD_POINT", "Got error while uncompressing: BLST_BAD_ENCODING")
    if b[0] & 0x40:
        if (b[0] & 0x3F) or any(b[1:]):
            raise KzgError("BAD_POINT", "Got error while uncompressing: BLST_BAD_ENCODING")
        return None
    x = int.from_bytes(bytes([b[0] & 0x1F]) + b[1:], "big")
    if x >= Q:
        raise KzgError("BAD_POINT", "Got error while uncompressing: BLST_BAD_ENCODING")
    y2 = (x * x * x + 4) % Q
    y = pow(y2, (Q + 1) // 4, Q)
    if y * y % Q != y2:
        raise KzgError("BAD_POINT", "Got error while uncompressing: BLST_POINT_NOT_ON_CURVE")
    if (y > HALF_Q) != bool(b[0] & 0x20):
        y = Q - y
    return (x, y)


# --------------------------------------------------------------------------
# Fq2 / G2 — only what verify_proof needs (src/curves.rs:185-343)
# --------------------------------------------------------------------------
def f2_add(a, b): return ((a[0] + b[0]) % Q, (a[1] + b[1]) % Q)
def f2_sub(a, b): return ((a[0] - b[0]) % Q, (a[1] - b[1]) % Q)
def f2_neg(a): return ((-a[0]) % Q, (-a[1]) % Q)
def f2_mul(a, b):
    return ((a[0] * b[0] - a[1] * b[1]) % Q, (a[0] * b[1] + a[1] * b[0]) % Q)
def f2_inv(a):
    d = pow(a[0] * a[0] + a[1] * a[1], Q - 2, Q)
    return (a[0] * d % Q, (-a[1]) * d % Q)
def f2_muli(a, k): return (a[0] * k % Q, a[1] * k % Q)


G2 = (G2_X, G2_Y)
B2 = (4, 4)


def g2_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return f2_sub(f2_mul(y, y), f2_add(f2_mul(f2_mul(x, x), x), B2)) == (0, 0)


def g2_neg(P):
    return None if P is None else (P[0], f2_neg(P[1]))


def g2_add(P, Qp):
    if P is None:
        return Qp
    if Qp is None:
        return P
    x1, y1 = P
    x2, y2 = Qp
    if x1 == x2:
        if y1 == y2 and y1 != (0, 0):
            lam = f2_mul(f2_muli(f2_mul(x1, x1), 3), f2_inv(f2_muli(y1, 2)))
        else:
            return None
    else:
        lam = f2_mul(f2_sub(y2, y1), f2_inv(f2_sub(x2, x1)))
    x3 = f2_sub(f2_sub(f2_mul(lam, lam), x1), x2)
    y3 = f2_sub(f2_mul(lam, f2_sub(x1, x3)), y1)
    return (x3, y3)


def g2_mul(P, k: int):
    acc = None
    for bit in bin(k)[2:] if k else "":
        acc = g2_add(acc, acc)
        if bit == "1":
            acc = g2_add(acc, P)
    return acc


def g2_sub(P, Qp):
    """G2Point::sub (src/curves.rs:241-249)."""
    return g2_add(P, g2_neg(Qp))


def g2_from_scalar(a: int):
    """G2Point::from_scalar (src/curves.rs:225-236)."""
    return g2_mul(G2, a % R)


def g2_from_i128(a: int):
    """G2Point::from_i128 (src/curves.rs:203-220)."""
    P = g2_mul(G2, abs(a))
    return g2_neg(P) if a < 0 else P


def g2_compress(P) -> bytes:
    """blst_p2_compress (src/curves.rs:260-271): 96 bytes = x.c1 || x.c0, flags in
    byte 0, sign from y.c1 unless zero, then y.c0."""
    if P is None:
        return bytes([0xC0]) + bytes(95)
    (x0, x1), (y0, y1) = P
    out = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    out[0] |= 0x80
    sign = (y1 > HALF_Q) if y1 != 0 else (y0 > HALF_Q)
    if sign:
        out[0] |= 0x20
    return bytes(out)


# --------------------------------------------------------------------------
# Pairing (equality oracle only) — bilinear_map, src/curves.rs:355-371.
# Fq12 = Fq[w]/(w^12 - 2 w^6 + 2); i = w^6 - 1.  Only lhs == rhs is ever used
# (src/polynomial.rs:293), so any correct non-degenerate pairing will do.
# --------------------------------------------------------------------------
def _f12_mul(a, b):
    t = [0] * 23
    for i, ai in enumerate(a):
        if ai:
            for j, bj in enumerate(b):
                t[i + j] += ai * bj
    # w^12 = 2 w^6 - 2
    for k in range(22, 11, -1):
        c = t[k]
        if c:
            t[k - 6] += 2 * c
            t[k - 12] -= 2 * c
    return [v % Q for v in t[:12]]


def _f12_one():
    return [1] + [0] * 11


def _f12_inv(a):
    # extended Euclid over Fq[w]
    def deg(p):
        d = len(p) - 1
        while d >= 0 and p[d] % Q == 0:
            d -= 1
        return d

    mod = [2, 0, 0, 0, 0, 0, Q - 2, 0, 0, 0, 0, 0, 1]
    lm, hm = [1] + [0] * 12, [0] * 13
    low, high = list(a) + [0], mod
    while deg(low) > 0:
        dl, dh = deg(low), deg(high)
        # r = high / low
        r = [0] * 13
        temp = list(high)
        inv_lead = pow(low[dl], Q - 2, Q)
        for i in range(dh - dl, -1, -1):
            c = temp[dl + i] * inv_lead % Q
            r[i] = c
            if c:
                for j in range(dl + 1):
                    temp[i + j] = (temp[i + j] - c * low[j]) % Q
        nm = list(hm)
        for i in range(13):
            if lm[i]:
                for j in range(13 - i):
                    if r[j]:
                        nm[i + j] = (nm[i + j] - lm[i] * r[j]) % Q
        lm, low, hm, high = nm, temp, lm, low
    inv0 = pow(low[0], Q - 2, Q)
    return [c * inv0 % Q for c in lm[:12]]


def _f12_pow(a, e: int):
    res = _f12_one()
    base = a
    while e:
        if e & 1:
            res = _f12_mul(res, base)
        base = _f12_mul(base, base)
        e >>= 1
    return res


def _f12_from_fq(x):
    return [x % Q] + [0] * 11


def _f12_from_f2(a):
    # (a0 + a1 i), i = w^6 - 1
    v = [0] * 12
    v[0] = (a[0] - a[1]) % Q
    v[6] = a[1] % Q
    return v


_W2_INV = None
_W3_INV = None


def _twist(P):
    """Map a point of E2 (the twist) to E(Fq12): (x/w^2, y/w^3)."""
    global _W2_INV, _W3_INV
    if _W2_INV is None:
        w2 = [0, 0, 1] + [0] * 9
        w3 = [0, 0, 0, 1] + [0] * 8
        _W2_INV, _W3_INV = _f12_inv(w2), _f12_inv(w3)
    x, y = P
    return (_f12_mul(_f12_from_f2(x), _W2_INV), _f12_mul(_f12_from_f2(y), _W3_INV))


def _f12_add(a, b): return [(x + y) % Q for x, y in zip(a, b)]
def _f12_sub(a, b): return [(x - y) % Q for x, y in zip(a, b)]


def _line(P1, P2, T):
    """Line through P1,P2 (points on E(Fq12)) evaluated at T."""
    x1, y1 = P1
    x2, y2 = P2
    xt, yt = T
    if x1 != x2:
        m = _f12_mul(_f12_sub(y2, y1), _f12_inv(_f12_sub(x2, x1)))
        return _f12_sub(_f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))
    if y1 == y2:
        three_x2 = _f12_mul(_f12_from_fq(3), _f12_mul(x1, x1))
        m = _f12_mul(three_x2, _f12_inv(_f12_add(y1, y1)))
        return _f12_sub(_f12_mul(m, _f12_sub(xt, x1)), _f12_sub(yt, y1))
    return _f12_sub(xt, x1)


def _e12_add(P1, P2):
    x1, y1 = P1
    x2, y2 = P2
    if x1 == x2 and y1 == y2:
        m = _f12_mul(_f12_mul(_f12_from_fq(3), _f12_mul(x1, x1)), _f12_inv(_f12_add(y1, y1)))
    else:
        m = _f12_mul(_f12_sub(y2, y1), _f12_inv(_f12_sub(x2, x1)))
    x3 = _f12_sub(_f12_sub(_f12_mul(m, m), x1), x2)
    y3 = _f12_sub(_f12_mul(m, _f12_sub(x1, x3)), y1)
    return (x3, y3)


_FINAL_EXP = (Q ** 12 - 1) // R


def pairing(P1, P2):
    """bilinear_map(p1, p2) up to inversion (sign of x ignored) — equality oracle
    (src/curves.rs:355-371: miller_loop(Q in G2, P in G1) then final_exp)."""
    if P1 is None or P2 is None:
        return _f12_one()
    Qt = _twist(P2)
    Pt = (_f12_from_fq(P1[0]), _f12_from_fq(P1[1]))
    Rr = Qt
    f = _f12_one()
    for bit in bin(BLS_X)[3:]:
        f = _f12_mul(_f12_mul(f, f), _line(Rr, Rr, Pt))
        Rr = _e12_add(Rr, Rr)
        if bit == "1":
            f = _f12_mul(f, _line(Rr, Qt, Pt))
            Rr = _e12_add(Rr, Qt)
    return _f12_pow(f, _FINAL_EXP)


# --------------------------------------------------------------------------
# Trusted setup — src/trusted_setup.rs
# --------------------------------------------------------------------------
class SetupArtifact:
    """SetupArtifact{g1,g2} (src/trusted_setup.rs:31-35).  g2 is computed lazily
    (only setup[1].g2 is ever consumed, src/polynomial.rs:284-286)."""

    __slots__ = ("g1", "_g2", "_spow")

    def __init__(self, g1, spow):
        self.g1 = g1
        self._spow = spow
        self._g2 = False

    @property
    def g2(self):
        if self._g2 is False:
            self._g2 = g2_mul(G2, self._spow)
        return self._g2


def setup_generate(secret_be: bytes, n: int, fast: bool = True):
    """SetupArtifactsGenerator::new(secret).take(n).collect() (src/trusted_setup.rs:20-28,40-78).
    Item 0 is the generator pair; item i>0 is ([s^i]G1, [s^i]G2), s = secret mod r.
    fast=True computes [s^i]G1 from the running Fr power by one scalar multiplication,
    exactly as the reference does."""
    s = fr_from_be_bytes(secret_be)
    out = []
    cur = 1
    for i in range(n):
        if i == 0:
            out.append(SetupArtifact(G1, 1))
            continue
        cur = cur * s % R
        out.append(SetupArtifact(g1_mul(G1, cur), cur))
    return out


# --------------------------------------------------------------------------
# Polynomial / Evaluation — src/polynomial.rs
# --------------------------------------------------------------------------
class Polynomial:
    """Polynomial{coefficients} ascending degree (src/polynomial.rs:9-12)."""

    def __init__(self, coefficients, _raw=False):
        self.coefficients = list(coefficients)

    @classmethod
    def try_from_scalars(cls, values):
        """TryFrom<Vec<Scalar>>: trailing zeros trimmed, all-zero input keeps ONE
        zero coefficient, empty input stays empty (src/polynomial.rs:55-75)."""
        values = [v % R for v in values]
        if len(values) > 0xFFFFFFFF:
            raise KzgError("BAD_LENGTH", "Too many coefficients for polynomial")
        last = 0
        for i, v in enumerate(values):
            if v != 0:
                last = i
        return cls(values[: last + 1])

    @classmethod
    def try_from_i128(cls, values):
        """TryFrom<Vec<i128>> (src/polynomial.rs:14-36)."""
        return cls.try_from_scalars([fr_from_i128(v) for v in values])

    @classmethod
    def from_scalar(cls, value):
        """From<Scalar>: zero gives the EMPTY polynomial (src/polynomial.rs:78-89)."""
        value %= R
        return cls([value] if value != 0 else [])

    def degree(self) -> int:
        """src/polynomial.rs:95-100."""
        return 0 if not self.coefficients else len(self.coefficients) - 1

    def evaluate(self, x: int):
        """Polynomial::evaluate — sum c_i * pow(x, i) (src/polynomial.rs:112-123)."""
        x %= R
        acc = 0
        for deg, c in enumerate(self.coefficients):
            acc = (acc + c * pow(x, deg, R)) % R
        return Evaluation(x, acc)

    def evaluate_ref_pow(self, x: int, counter=None):
        """Same as evaluate() but through the reference's bespoke pow (src/scalar.rs:122-187)."""
        x %= R
        acc = 0
        for deg, c in enumerate(self.coefficients):
            acc = (acc + c * fr_pow_ref(x, deg, counter)) % R
        return Evaluation(x, acc)

    def sub(self, other: "Polynomial") -> "Polynomial":
        """Polynomial::sub (src/polynomial.rs:128-145)."""
        a, b = self.coefficients, other.coefficients
        if len(a) > len(b):
            co = list(a)
            for i, rhs in enumerate(b):
                co[i] = (co[i] - rhs) % R
        else:
            co = [(-x) % R for x in b]
            for i, lhs in enumerate(a):
                co[i] = (lhs + co[i]) % R
        return Polynomial.try_from_scalars(co)

    def divide_by_root(self, root: int) -> "Polynomial":
        """Polynomial::divide_by_root — synthetic division with the remainder check
        (src/polynomial.rs:150-195)."""
        root %= R
        if not self.coefficients:
            return Polynomial([])
        hi = self.coefficients[-1]
        if len(self.coefficients) == 1:
            if hi == 0:
                return Polynomial([])
            raise KzgError("CONSTANT_POLY", "Unable to divide a constant polynomial")
        desc = [hi]
        degree = self.degree()
        for i in range(degree - 1, 0, -1):
            desc.append((self.coefficients[i] + root * desc[degree - i - 1]) % R)
        desc.reverse()
        rebuilt = (-(root * desc[0])) % R
        if rebuilt != self.coefficients[0]:
            raise KzgError(
                "REMAINDER_NONZERO",
                "[divide_by_root] Fail to divide the polynomial by a root, constant terms do not add up",
            )
        return Polynomial.try_from_scalars(desc)

    def commit(self, setup):
        """Polynomial::commit — naive sum of [c_i] setup[i].g1 (src/polynomial.rs:200-215)."""
        if self.degree() + 1 > len(setup):
            raise KzgError(
                "DEGREE_TOO_HIGH",
                "Setup does not allow for commitment generation of the polynomial. The polynomial degree is too high.",
            )
        acc = (1, 1, 0)
        for i, c in enumerate(self.coefficients):
            g1 = setup[i].g1 if isinstance(setup[i], SetupArtifact) else setup[i]
            acc = _jac_add(acc, _jac_mul(_to_jac(g1), c))
        return _from_jac(acc)


class Evaluation:
    """Evaluation{point,result} (src/polynomial.rs:249-253)."""

    def __init__(self, point: int, result: int):
        self.point = point % R
        self.result = result % R

    def generate_proof(self, polynomial: Polynomial, setup):
        """Evaluation::generate_proof (src/polynomial.rs:260-269)."""
        return (
            polynomial.sub(Polynomial.from_scalar(self.result))
            .divide_by_root(self.point)
            .commit(setup)
        )

    def verify_proof(self, proof, commitment, setup) -> bool:
        """Evaluation::verify_proof: e(proof,[s]2-[z]2) == e(C-[y]1, G2)
        (src/polynomial.rs:276-294).  Indexes setup[1] unguarded like the reference."""
        lhs = pairing(proof, g2_sub(setup[1].g2, g2_from_scalar(self.point)))
        rhs = pairing(g1_sub(commitment, g1_from_scalar(self.result)), g2_from_i128(1))
        return lhs == rhs


# --------------------------------------------------------------------------
# Known-secret shortcuts (valid because every benchmark setup comes from a known s)
# --------------------------------------------------------------------------
def horner(coeffs, x: int) -> int:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % R
    return acc


def quotient(coeffs, z: int):
    """(q, y): q(x) = (p(x)-p(z))/(x-z), y = p(z) — Horner intermediates
    (equivalent to src/polynomial.rs:168-179 after the :266 subtraction)."""
    n = len(coeffs)
    if n == 0:
        return [], 0
    q = [0] * (n - 1)
    t = coeffs[-1] % R
    for i in range(n - 2, -1, -1):
        q[i] = t
        t = (coeffs[i] + z * t) % R
    return q, t


def commit_known_secret(coeffs, s: int):
    """commit(p) == [p(s)] G1 when setup[i] = [s^i] G1."""
    return g1_mul(G1, horner(coeffs, s % R))


def proof_known_secret(coeffs, s: int, z: int):
    q, _ = quotient(coeffs, z % R)
    return g1_mul(G1, horner(q, s % R))


# --------------------------------------------------------------------------
# Deterministic inputs
# --------------------------------------------------------------------------
BENCH_SECRET_BE = bytes(range(32))  # benches/polynomial_commitment.rs:17-19


def bench_polynomial(degree: int):
    """c_i = 5^i + 10 (benches/polynomial_commitment.rs:10-15)."""
    return [(pow(5, i, R) + 10) % R for i in range(degree + 1)]


def bench_input_point(degree: int) -> int:
    """z = 5^d + 20 (benches/evaluation_proof.rs:25-27)."""
    return (pow(5, degree, R) + 20) % R


_M64 = (1 << 64) - 1


def _splitmix(z: int) -> int:
    z &= _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def scalar_stream(seed: int, k: int) -> int:
    """U(seed,k): counter-based ~uniform Fr element — identical in
    Python, C and CUDA."""
    v = 0
    for j in range(4):
        w = _splitmix(seed + (4 * k + j + 1) * 0x9E3779B97F4A7C15)
        v |= w << (64 * j)
    v &= (1 << 255) - 1
    if v >= R:
        v -= R
    return v
