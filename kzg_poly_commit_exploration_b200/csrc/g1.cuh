// G1 of BLS12-381 (y^2 = x^3 + 4 over Fq) in affine and XYZZ coordinates.
//
// Replaces the reference's blst_p1_* calls: blst_p1_add_or_double (src/curves.rs:71,82),
// blst_p1_cneg (:41,70), blst_p1_mult (:32,53,93; src/trusted_setup.rs:56), blst_p1_compress
// (:106), blst_p1_uncompress / _from_affine (:131,141), blst_p1_generator (:34,55).
//
// XYZZ: x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2, infinity <=> ZZ == 0.  All additions are COMPLETE:
// they absorb infinity operands, P == Q (doubling) and P == -Q exactly like
// blst_p1_add_or_double, because reference inputs hit those cases (equal setup points when
// s = 1, infinity bases when s = 0, cancelling buckets).
// Affine points store Montgomery coordinates; (0,0) — not on the curve — marks infinity.
#pragma once
#include "field.cuh"

namespace kzg {

struct G1Affine { Fq x, y; };           // 96 bytes, (0,0) = infinity
struct G1Xyzz { Fq x, y, zz, zzz; };    // 192 bytes

KZG_HD bool is_inf(const G1Affine& p) { return is_zero(p.x) && is_zero(p.y); }
KZG_HD bool is_inf(const G1Xyzz& p) { return is_zero(p.zz); }
KZG_HD void set_inf(G1Affine& p) { set_zero(p.x); set_zero(p.y); }
KZG_HD void set_inf(G1Xyzz& p) { set_zero(p.x); set_zero(p.y); set_zero(p.zz); set_zero(p.zzz); }

KZG_HD void g1_generator(G1Affine& g) {
#pragma unroll
  for (int i = 0; i < 12; i++) { g.x.v[i] = KZG_C(FQ_G1X)[i]; g.y.v[i] = KZG_C(FQ_G1Y)[i]; }
}

KZG_HD void from_affine(G1Xyzz& o, const G1Affine& p) {
  if (is_inf(p)) { set_inf(o); return; }
  o.x = p.x; o.y = p.y; set_one(o.zz); set_one(o.zzz);
}

// 2*(x,y) for an affine point (y != 0 on this curve: no 2-torsion in E(Fq) of odd order part;
// a y == 0 input would give ZZ = 0 = infinity, which is also the right answer).
KZG_HD void dbl_affine(G1Xyzz& o, const Fq& x, const Fq& y) {
  Fq u, v, w, s, m, t;
  dbl(u, y);            // U = 2y
  sqr(v, u);            // V = U^2
  mul(w, u, v);         // W = U*V
  mul(s, x, v);         // S = x*V
  sqr(m, x);            // x^2
  dbl(t, m); add(m, t, m);  // M = 3x^2
  sqr(o.x, m); sub(o.x, o.x, s); sub(o.x, o.x, s);   // X3 = M^2 - 2S
  sub(t, s, o.x); mul(t, m, t);                       // M*(S-X3)
  mul(u, w, y);                                       // W*y
  sub(o.y, t, u);
  o.zz = v; o.zzz = w;
}

// acc = 2*acc
KZG_HD void dbl(G1Xyzz& a) {
  if (is_inf(a)) return;
  Fq u, v, w, s, m, t;
  dbl(u, a.y);
  sqr(v, u);
  mul(w, u, v);
  mul(s, a.x, v);
  sqr(m, a.x);
  dbl(t, m); add(m, t, m);
  Fq x3;
  sqr(x3, m); sub(x3, x3, s); sub(x3, x3, s);
  sub(t, s, x3); mul(t, m, t);
  mul(u, w, a.y);
  sub(a.y, t, u);
  a.x = x3;
  mul(a.zz, v, a.zz);
  mul(a.zzz, w, a.zzz);
}

// acc += (px, py) affine, not infinity (caller filters the (0,0) marker).  8M + 2S.
KZG_HD void madd(G1Xyzz& a, const Fq& px, const Fq& py) {
  if (is_inf(a)) {
    a.x = px; a.y = py; set_one(a.zz); set_one(a.zzz);
    return;
  }
  Fq p, r, pp, ppp, q, t;
  mul(p, px, a.zz);  sub(p, p, a.x);     // P = U2 - X1
  mul(r, py, a.zzz); sub(r, r, a.y);     // R = S2 - Y1
  if (is_zero(p)) {
    if (is_zero(r)) dbl_affine(a, px, py); else set_inf(a);
    return;
  }
  sqr(pp, p);
  mul(ppp, p, pp);
  mul(q, a.x, pp);
  sqr(a.x, r); sub(a.x, a.x, ppp); sub(a.x, a.x, q); sub(a.x, a.x, q);   // X3
  sub(t, q, a.x); mul(t, r, t);
  mul(q, a.y, ppp);
  sub(a.y, t, q);                                                         // Y3
  mul(a.zz, a.zz, pp);
  mul(a.zzz, a.zzz, ppp);
}

KZG_HD void madd(G1Xyzz& a, const G1Affine& p, bool negate) {
  if (is_inf(p)) return;
  // branch-free conditional negation: both signs must run the SAME madd instructions, otherwise a warp
  // splits into a "+" half and a "-" half that execute two inlined copies of the addition
  Fq y; neg(y, p.y);
#pragma unroll
  for (int i = 0; i < 12; i++) y.v[i] = negate ? y.v[i] : p.y.v[i];
  madd(a, p.x, y);
}

// acc += b.  12M + 2S.
KZG_HD void add(G1Xyzz& a, const G1Xyzz& b) {
  if (is_inf(b)) return;
  if (is_inf(a)) { a = b; return; }
  Fq u1, u2, s1, s2, p, r, pp, ppp, q, t;
  mul(u1, a.x, b.zz);
  mul(u2, b.x, a.zz);
  mul(s1, a.y, b.zzz);
  mul(s2, b.y, a.zzz);
  sub(p, u2, u1);
  sub(r, s2, s1);
  if (is_zero(p)) {
    if (is_zero(r)) dbl(a); else set_inf(a);
    return;
  }
  sqr(pp, p);
  mul(ppp, p, pp);
  mul(q, u1, pp);
  sqr(a.x, r); sub(a.x, a.x, ppp); sub(a.x, a.x, q); sub(a.x, a.x, q);
  sub(t, q, a.x); mul(t, r, t);
  mul(s1, s1, ppp);
  sub(a.y, t, s1);
  mul(a.zz, a.zz, b.zz);   mul(a.zz, a.zz, pp);
  mul(a.zzz, a.zzz, b.zzz); mul(a.zzz, a.zzz, ppp);
}

KZG_HD void neg(G1Xyzz& a) { neg(a.y, a.y); }

// XYZZ -> affine with a caller-supplied inverse of ZZZ (batch inversion) : 1/ZZ = ZZ^2 * (1/ZZZ)^2.
KZG_HD void to_affine_with_inv(G1Affine& o, const G1Xyzz& a, const Fq& zzz_inv) {
  if (is_inf(a)) { set_inf(o); return; }
  Fq w2, z2, zi;
  sqr(w2, zzz_inv);
  sqr(z2, a.zz);
  mul(zi, z2, w2);          // 1/ZZ
  mul(o.x, a.x, zi);
  mul(o.y, a.y, zzz_inv);
}
KZG_HD void to_affine(G1Affine& o, const G1Xyzz& a) {
  if (is_inf(a)) { set_inf(o); return; }
  Fq w; inv_gcd(w, a.zzz);
  to_affine_with_inv(o, a, w);
}

// [k]P, k a canonical 256-bit integer in 8 limbs — plain double-and-add, used for single
// element API calls (G1Point::mult / from_scalar, src/curves.rs:50-61,90-96) and table seeds.
KZG_HD void scalar_mul(G1Xyzz& o, const G1Affine& p, const uint32_t* k, int nbits) {
  G1Xyzz acc; set_inf(acc);
  for (int i = nbits - 1; i >= 0; i--) {
    dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) madd(acc, p, false);
  }
  o = acc;
}

// ZCash 48-byte compression of an affine point (blst_p1_compress, src/curves.rs:99-110).
KZG_HD void compress(uint8_t* out, const G1Affine& p) {
  if (is_inf(p)) {
    out[0] = 0xC0;
    for (int i = 1; i < 48; i++) out[i] = 0;
    return;
  }
  Fq x, y;
  from_mont(x, p.x); from_mont(y, p.y);
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t w = x.v[11 - i];
    out[4 * i] = (uint8_t)(w >> 24); out[4 * i + 1] = (uint8_t)(w >> 16);
    out[4 * i + 2] = (uint8_t)(w >> 8); out[4 * i + 3] = (uint8_t)w;
  }
  out[0] |= 0x80;
  if (fq_canon_gt_half(y)) out[0] |= 0x20;
}

// 48 bytes -> affine (blst_p1_uncompress semantics used by bytes_to_blst_p1,
// src/curves.rs:119-144): flag checks, x < q, sqrt, on-curve; NO subgroup check.
// Returns 0 on success, nonzero on a malformed / off-curve encoding.
KZG_HD int uncompress(G1Affine& o, const uint8_t* in) {
  uint8_t b0 = in[0];
  if (!(b0 & 0x80)) return 1;
  if (b0 & 0x40) {
    if (b0 & 0x3F) return 1;
    for (int i = 1; i < 48; i++) if (in[i]) return 1;
    set_inf(o);
    return 0;
  }
  Fq x;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t w = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) | ((uint32_t)in[4 * i + 2] << 8) | in[4 * i + 3];
    if (i == 0) w &= 0x1FFFFFFFu;
    x.v[11 - i] = w;
  }
  if (!fq_canon_lt_p(x)) return 1;
  Fq xm, y2, y, t;
  to_mont(xm, x);
  sqr(y2, xm); mul(y2, y2, xm);
  Fq b4;
#pragma unroll
  for (int i = 0; i < 12; i++) b4.v[i] = KZG_C(FQ_B4)[i];
  add(y2, y2, b4);
  pow_limbs<Fq, 12>(y, y2, KZG_C(FQ_SQRT_EXP));
  sqr(t, y);
  if (!eq(t, y2)) return 2;
  Fq yc; from_mont(yc, y);
  bool big = fq_canon_gt_half(yc);
  if (big != ((b0 & 0x20) != 0)) neg(y, y);
  o.x = xm; o.y = y;
  return 0;
}

}  // namespace kzg
