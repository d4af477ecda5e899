// Fq2 = Fq[i]/(i^2 + 1) and G2 of BLS12-381 (the twist y^2 = x^3 + 4(1+i) over Fq2), host + device.
//
// Replaces the reference's blst_p2_* calls: blst_p2_mult / blst_p2_generator (src/curves.rs:203-236,
// src/trusted_setup.rs:64-72 — the G2 half of every SetupArtifact), blst_p2_cneg + blst_p2_add_or_double
// (G2Point::sub, src/curves.rs:241-249), blst_p2_compress (:260-271), blst_p2_uncompress / _from_affine
// (:279-297).  Same XYZZ formulas as g1.cuh (the curve shape is the same, a = 0), instantiated over Fq2;
// all additions are complete.  Affine (0,0) marks infinity.
#pragma once
#include "g1.cuh"

namespace kzg {

struct Fq2 { Fq c0, c1; };               // c0 + c1*i, 96 bytes

KZG_HD bool is_zero(const Fq2& a) { return is_zero(a.c0) && is_zero(a.c1); }
KZG_HD bool eq(const Fq2& a, const Fq2& b) { return eq(a.c0, b.c0) && eq(a.c1, b.c1); }
KZG_HD void set_zero(Fq2& a) { set_zero(a.c0); set_zero(a.c1); }
KZG_HD void set_one(Fq2& a) { set_one(a.c0); set_zero(a.c1); }
KZG_HD void add(Fq2& o, const Fq2& a, const Fq2& b) { add(o.c0, a.c0, b.c0); add(o.c1, a.c1, b.c1); }
KZG_HD void sub(Fq2& o, const Fq2& a, const Fq2& b) { sub(o.c0, a.c0, b.c0); sub(o.c1, a.c1, b.c1); }
KZG_HD void neg(Fq2& o, const Fq2& a) { neg(o.c0, a.c0); neg(o.c1, a.c1); }
KZG_HD void dbl(Fq2& o, const Fq2& a) { add(o.c0, a.c0, a.c0); add(o.c1, a.c1, a.c1); }
KZG_HD void conj(Fq2& o, const Fq2& a) { o.c0 = a.c0; neg(o.c1, a.c1); }
// (a0 + a1 i)(b0 + b1 i) = (a0 b0 - a1 b1) + ((a0 + a1)(b0 + b1) - a0 b0 - a1 b1) i : 3 Fq products
KZG_HD void mul(Fq2& o, const Fq2& a, const Fq2& b) {
  Fq v0, v1, s, t;
  mul(v0, a.c0, b.c0);
  mul(v1, a.c1, b.c1);
  add(s, a.c0, a.c1);
  add(t, b.c0, b.c1);
  mul(s, s, t);
  sub(s, s, v0);
  sub(o.c1, s, v1);
  sub(o.c0, v0, v1);
}
// (a0 + a1 i)^2 = (a0 + a1)(a0 - a1) + 2 a0 a1 i : 2 Fq products
KZG_HD void sqr(Fq2& o, const Fq2& a) {
  Fq s, d, p;
  add(s, a.c0, a.c1);
  sub(d, a.c0, a.c1);
  mul(p, a.c0, a.c1);
  mul(o.c0, s, d);
  add(o.c1, p, p);
}
KZG_HD void mul_fq(Fq2& o, const Fq2& a, const Fq& k) { mul(o.c0, a.c0, k); mul(o.c1, a.c1, k); }
// 1/(a0 + a1 i) = (a0 - a1 i)/(a0^2 + a1^2)
KZG_HD void inv(Fq2& o, const Fq2& a) {
  Fq n, t;
  sqr(n, a.c0); sqr(t, a.c1); add(n, n, t);
  inv_gcd(t, n);
  mul(o.c0, a.c0, t);
  mul(t, a.c1, t);
  neg(o.c1, t);
}

struct G2Affine { Fq2 x, y; };            // 192 bytes, (0,0) = infinity
struct G2Xyzz { Fq2 x, y, zz, zzz; };     // 384 bytes

KZG_HD bool is_inf(const G2Affine& p) { return is_zero(p.x) && is_zero(p.y); }
KZG_HD bool is_inf(const G2Xyzz& p) { return is_zero(p.zz); }
KZG_HD void set_inf(G2Affine& p) { set_zero(p.x); set_zero(p.y); }
KZG_HD void set_inf(G2Xyzz& p) { set_zero(p.x); set_zero(p.y); set_zero(p.zz); set_zero(p.zzz); }

KZG_HD void g2_generator(G2Affine& g) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    g.x.c0.v[i] = KZG_C(FQ_G2X0)[i]; g.x.c1.v[i] = KZG_C(FQ_G2X1)[i];
    g.y.c0.v[i] = KZG_C(FQ_G2Y0)[i]; g.y.c1.v[i] = KZG_C(FQ_G2Y1)[i];
  }
}

KZG_HD void from_affine(G2Xyzz& o, const G2Affine& p) {
  if (is_inf(p)) { set_inf(o); return; }
  o.x = p.x; o.y = p.y; set_one(o.zz); set_one(o.zzz);
}

KZG_HD void dbl_affine(G2Xyzz& o, const Fq2& x, const Fq2& y) {
  Fq2 u, v, w, s, m, t;
  dbl(u, y);
  sqr(v, u);
  mul(w, u, v);
  mul(s, x, v);
  sqr(m, x);
  dbl(t, m); add(m, t, m);
  sqr(o.x, m); sub(o.x, o.x, s); sub(o.x, o.x, s);
  sub(t, s, o.x); mul(t, m, t);
  mul(u, w, y);
  sub(o.y, t, u);
  o.zz = v; o.zzz = w;
}

KZG_HD void dbl(G2Xyzz& a) {
  if (is_inf(a)) return;
  Fq2 u, v, w, s, m, t, x3;
  dbl(u, a.y);
  sqr(v, u);
  mul(w, u, v);
  mul(s, a.x, v);
  sqr(m, a.x);
  dbl(t, m); add(m, t, m);
  sqr(x3, m); sub(x3, x3, s); sub(x3, x3, s);
  sub(t, s, x3); mul(t, m, t);
  mul(u, w, a.y);
  sub(a.y, t, u);
  a.x = x3;
  mul(a.zz, v, a.zz);
  mul(a.zzz, w, a.zzz);
}

// acc += (px, py) affine, not infinity
KZG_HD void madd(G2Xyzz& a, const Fq2& px, const Fq2& py) {
  if (is_inf(a)) { a.x = px; a.y = py; set_one(a.zz); set_one(a.zzz); return; }
  Fq2 p, r, pp, ppp, q, t;
  mul(p, px, a.zz);  sub(p, p, a.x);
  mul(r, py, a.zzz); sub(r, r, a.y);
  if (is_zero(p)) {
    if (is_zero(r)) dbl_affine(a, px, py); else set_inf(a);
    return;
  }
  sqr(pp, p);
  mul(ppp, p, pp);
  mul(q, a.x, pp);
  sqr(a.x, r); sub(a.x, a.x, ppp); sub(a.x, a.x, q); sub(a.x, a.x, q);
  sub(t, q, a.x); mul(t, r, t);
  mul(q, a.y, ppp);
  sub(a.y, t, q);
  mul(a.zz, a.zz, pp);
  mul(a.zzz, a.zzz, ppp);
}
KZG_HD void madd(G2Xyzz& a, const G2Affine& p, bool negate) {
  if (is_inf(p)) return;
  Fq2 y = p.y;
  if (negate) neg(y, p.y);
  madd(a, p.x, y);
}

KZG_HD void add(G2Xyzz& a, const G2Xyzz& b) {
  if (is_inf(b)) return;
  if (is_inf(a)) { a = b; return; }
  Fq2 u1, u2, s1, s2, p, r, pp, ppp, q, t;
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

KZG_HD void to_affine_with_inv(G2Affine& o, const G2Xyzz& a, const Fq2& zzz_inv) {
  if (is_inf(a)) { set_inf(o); return; }
  Fq2 w2, z2, zi;
  sqr(w2, zzz_inv);
  sqr(z2, a.zz);
  mul(zi, z2, w2);
  mul(o.x, a.x, zi);
  mul(o.y, a.y, zzz_inv);
}
KZG_HD void to_affine(G2Affine& o, const G2Xyzz& a) {
  if (is_inf(a)) { set_inf(o); return; }
  Fq2 w; inv(w, a.zzz);
  to_affine_with_inv(o, a, w);
}

// [k]P, k a canonical integer in 32-bit limbs — double-and-add (G2Point::from_scalar / from_i128)
KZG_HD void scalar_mul(G2Xyzz& o, const G2Affine& p, const uint32_t* k, int nbits) {
  G2Xyzz acc; set_inf(acc);
  for (int i = nbits - 1; i >= 0; i--) {
    dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) madd(acc, p, false);
  }
  o = acc;
}

KZG_HD void fq_store_be(uint8_t* out, const Fq& canon) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t w = canon.v[11 - i];
    out[4 * i] = (uint8_t)(w >> 24); out[4 * i + 1] = (uint8_t)(w >> 16);
    out[4 * i + 2] = (uint8_t)(w >> 8); out[4 * i + 3] = (uint8_t)w;
  }
}
KZG_HD void fq_load_be(Fq& canon, const uint8_t* in, uint32_t top_mask) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint32_t w = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) | ((uint32_t)in[4 * i + 2] << 8) | in[4 * i + 3];
    if (i == 0) w &= top_mask;
    canon.v[11 - i] = w;
  }
}

// ZCash 96-byte compression (blst_p2_compress, src/curves.rs:260-271): x.c1 || x.c0 big-endian, flags in
// byte 0, sign bit from y.c1 unless it is zero, then from y.c0.
KZG_HD void compress(uint8_t* out, const G2Affine& p) {
  if (is_inf(p)) {
    out[0] = 0xC0;
    for (int i = 1; i < 96; i++) out[i] = 0;
    return;
  }
  Fq x0, x1, y0, y1;
  from_mont(x0, p.x.c0); from_mont(x1, p.x.c1); from_mont(y0, p.y.c0); from_mont(y1, p.y.c1);
  fq_store_be(out, x1);
  fq_store_be(out + 48, x0);
  out[0] |= 0x80;
  bool sign = is_zero(y1) ? fq_canon_gt_half(y0) : fq_canon_gt_half(y1);
  if (sign) out[0] |= 0x20;
}

// sqrt in Fq2 for q = 3 mod 4 (Adj & Rodriguez-Henriquez, "Square root computation over even extension
// fields", Alg. 9).  Returns false when a is not a square.
KZG_HD bool sqrt_fq2(Fq2& o, const Fq2& a) {
  if (is_zero(a)) { set_zero(o); return true; }
  Fq2 a1, alpha, x0, t, one, mone;
  pow_limbs<Fq2, 12>(a1, a, KZG_C(FQ_SQRT2_EXP1));   // a^((q-3)/4)
  mul(x0, a1, a);
  mul(alpha, a1, x0);                                // a^((q-1)/2)
  set_one(one); neg(mone, one);
  conj(t, alpha); mul(t, t, alpha);                  // alpha^(q+1)
  if (eq(t, mone)) return false;
  if (eq(alpha, mone)) {                             // x = i * x0
    neg(o.c0, x0.c1); o.c1 = x0.c0;
  } else {
    add(t, one, alpha);
    Fq2 b; pow_limbs<Fq2, 12>(b, t, KZG_C(FQ_SQRT2_EXP2));   // (1 + alpha)^((q-1)/2)
    mul(o, b, x0);
  }
  sqr(t, o);
  return eq(t, a);
}

// 96 bytes -> affine (blst_p2_uncompress semantics used by bytes_to_blst_p2, src/curves.rs:279-297): flag
// checks, coordinates < q, sqrt, on-curve; NO subgroup check.  0 on success.
KZG_HD int uncompress(G2Affine& o, const uint8_t* in) {
  uint8_t b0 = in[0];
  if (!(b0 & 0x80)) return 1;
  if (b0 & 0x40) {
    if (b0 & 0x3F) return 1;
    for (int i = 1; i < 96; i++) if (in[i]) return 1;
    set_inf(o);
    return 0;
  }
  Fq x1c, x0c;
  fq_load_be(x1c, in, 0x1FFFFFFFu);
  fq_load_be(x0c, in + 48, 0xFFFFFFFFu);
  if (!fq_canon_lt_p(x1c) || !fq_canon_lt_p(x0c)) return 1;
  Fq2 x, y2, y, b;
  to_mont(x.c0, x0c); to_mont(x.c1, x1c);
  sqr(y2, x); mul(y2, y2, x);
#pragma unroll
  for (int i = 0; i < 12; i++) { b.c0.v[i] = KZG_C(FQ_B4)[i]; b.c1.v[i] = KZG_C(FQ_B4)[i]; }   // 4 + 4i
  add(y2, y2, b);
  if (!sqrt_fq2(y, y2)) return 2;
  Fq y0c, y1c;
  from_mont(y0c, y.c0); from_mont(y1c, y.c1);
  bool sign = is_zero(y1c) ? fq_canon_gt_half(y0c) : fq_canon_gt_half(y1c);
  if (sign != ((b0 & 0x20) != 0)) neg(y, y);
  o.x = x; o.y = y;
  return 0;
}

}  // namespace kzg
