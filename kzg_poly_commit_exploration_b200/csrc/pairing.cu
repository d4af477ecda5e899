// Host-side G2 helpers and the pairing check of Evaluation::verify_proof (src/polynomial.rs:276-294,
// bilinear_map src/curves.rs:355-371, G2Point::{from_scalar, from_i128, sub} :203-249).
//
// Two pairings whatever the polynomial degree: O(1) work, so it stays on the host; it
// is here because the reference gets it from blst and this library must stand on its own.  Only EQUALITY of
// pairings is ever consumed (src/polynomial.rs:293), so the check is done as one product of Miller loops
// followed by one final exponentiation:  e(A1,B1) == e(A2,B2)  <=>  FE(ML(A1,B1) * ML(-A2,B2)) == 1.
//
// Fq12 is kept flat: Fq[w]/(w^12 - 2 w^6 + 2), with Fq2 embedded by i = w^6 - 1 (so w^6 = 1 + i, the sextic
// non-residue of the tower).  The twist isomorphism is (x', y') -> (x'/w^2, y'/w^3); G2 arithmetic of the
// Miller loop runs on the twist in affine Fq2 coordinates, and each line is scaled by w^3 — an element whose
// (q^6-1)(q^6+1)/r-th power is 1, so the scaling disappears in the final exponentiation.
#include "../../include/kzg_b200.h"
#include "g2.cuh"

using namespace kzg;

namespace {

struct Fq12 { Fq c[12]; };

void f12_one(Fq12& a) { for (int i = 0; i < 12; i++) set_zero(a.c[i]); set_one(a.c[0]); }
bool f12_is_one(const Fq12& a) {
  Fq one; set_one(one);
  if (!eq(a.c[0], one)) return false;
  for (int i = 1; i < 12; i++) if (!is_zero(a.c[i])) return false;
  return true;
}

void f12_mul(Fq12& o, const Fq12& a, const Fq12& b) {
  Fq t[23];
  for (int k = 0; k < 23; k++) set_zero(t[k]);
  for (int i = 0; i < 12; i++) {
    if (is_zero(a.c[i])) continue;
    for (int j = 0; j < 12; j++) {
      if (is_zero(b.c[j])) continue;
      Fq p; mul(p, a.c[i], b.c[j]);
      add(t[i + j], t[i + j], p);
    }
  }
  for (int k = 22; k >= 12; k--) {       // w^12 = 2 w^6 - 2
    Fq d; add(d, t[k], t[k]);
    add(t[k - 6], t[k - 6], d);
    sub(t[k - 12], t[k - 12], d);
  }
  for (int k = 0; k < 12; k++) o.c[k] = t[k];
}

// w -> -w: the Frobenius of order 2 over Fq6 (x -> x^(q^6))
void f12_conj6(Fq12& o, const Fq12& a) {
  for (int k = 0; k < 12; k++) { if (k & 1) neg(o.c[k], a.c[k]); else o.c[k] = a.c[k]; }
}

// inverse by solving (multiplication-by-a matrix) x = 1 over Fq
bool f12_inv(Fq12& o, const Fq12& a) {
  Fq m[12][13];
  Fq12 col, wj;
  for (int j = 0; j < 12; j++) {
    for (int k = 0; k < 12; k++) set_zero(wj.c[k]);
    set_one(wj.c[j]);
    f12_mul(col, a, wj);
    for (int i = 0; i < 12; i++) m[i][j] = col.c[i];
  }
  for (int i = 0; i < 12; i++) set_zero(m[i][12]);
  set_one(m[0][12]);
  for (int c = 0; c < 12; c++) {
    int piv = -1;
    for (int r = c; r < 12; r++) if (!is_zero(m[r][c])) { piv = r; break; }
    if (piv < 0) return false;
    if (piv != c) for (int k = 0; k < 13; k++) { Fq t = m[c][k]; m[c][k] = m[piv][k]; m[piv][k] = t; }
    Fq pi; inv_gcd(pi, m[c][c]);
    for (int k = c; k < 13; k++) mul(m[c][k], m[c][k], pi);
    for (int r = 0; r < 12; r++) {
      if (r == c || is_zero(m[r][c])) continue;
      Fq f = m[r][c];
      for (int k = c; k < 13; k++) { Fq t; mul(t, f, m[c][k]); sub(m[r][k], m[r][k], t); }
    }
  }
  for (int i = 0; i < 12; i++) o.c[i] = m[i][12];
  return true;
}

// place c (in Fq2) * w^k into a flat element: c0 + c1 i = (c0 - c1) + c1 w^6
void f12_add_fq2_at(Fq12& a, const Fq2& c, int k) {
  Fq d; sub(d, c.c0, c.c1);
  add(a.c[k], a.c[k], d);
  add(a.c[k + 6], a.c[k + 6], c.c1);
}

// line through T and U on the twist (slope lam), evaluated at P = (xp, yp), times w^3:
//   yp w^3 - (lam xp) w^2 + (lam x_T - y_T)
void line_eval(Fq12& l, const Fq2& lam, const G2Affine& T, const Fq& xp, const Fq& yp) {
  for (int k = 0; k < 12; k++) set_zero(l.c[k]);
  Fq2 c0, c2;
  mul(c0, lam, T.x); sub(c0, c0, T.y);
  mul_fq(c2, lam, xp); neg(c2, c2);
  f12_add_fq2_at(l, c0, 0);
  f12_add_fq2_at(l, c2, 2);
  add(l.c[3], l.c[3], yp);
}

// one Miller step: f <- f * line(T, U)(P), T <- T + U (affine, on the twist).  U may equal T (doubling).
void miller_step(Fq12& f, G2Affine& T, const G2Affine& U, const Fq& xp, const Fq& yp) {
  if (is_inf(T)) { T = U; return; }
  if (is_inf(U)) return;
  Fq2 lam, num, den;
  Fq12 l;
  if (eq(T.x, U.x)) {
    Fq2 ny; neg(ny, U.y);
    if (!eq(T.y, U.y) || is_zero(T.y)) {
      // vertical line x - x_T, times w^2:  xp w^2 - x_T
      (void)ny;
      for (int k = 0; k < 12; k++) set_zero(l.c[k]);
      Fq2 nx; neg(nx, T.x);
      f12_add_fq2_at(l, nx, 0);
      add(l.c[2], l.c[2], xp);
      Fq12 t; f12_mul(t, f, l); f = t;
      set_inf(T);
      return;
    }
    sqr(num, T.x); Fq2 t2; dbl(t2, num); add(num, num, t2);   // 3 x^2
    dbl(den, T.y);
  } else {
    sub(num, U.y, T.y);
    sub(den, U.x, T.x);
  }
  Fq2 di; inv(di, den);
  mul(lam, num, di);
  line_eval(l, lam, T, xp, yp);
  Fq12 t; f12_mul(t, f, l); f = t;
  Fq2 x3, y3;
  sqr(x3, lam); sub(x3, x3, T.x); sub(x3, x3, U.x);
  sub(y3, T.x, x3); mul(y3, lam, y3); sub(y3, y3, T.y);
  T.x = x3; T.y = y3;
}

// f *= MillerLoop(Q, P) for the ate pairing with |x| = 0xd201000000010000 (the sign of x only conjugates the
// result, which does not change an equality test)
void miller_loop(Fq12& f_io, const G1Affine& P, const G2Affine& Qt) {
  if (is_inf(P) || is_inf(Qt)) return;
  const uint64_t x = 0xd201000000010000ull;
  Fq12 f; f12_one(f);
  G2Affine T = Qt;
  for (int b = 62; b >= 0; b--) {
    Fq12 s; f12_mul(s, f, f); f = s;
    G2Affine Tc = T;
    miller_step(f, T, Tc, P.x, P.y);
    if ((x >> b) & 1) miller_step(f, T, Qt, P.x, P.y);
  }
  Fq12 t; f12_mul(t, f_io, f); f_io = t;
}

bool final_exp_is_one(const Fq12& f) {
  Fq12 fi, fc, g;
  if (!f12_inv(fi, f)) return false;
  f12_conj6(fc, f);
  f12_mul(g, fc, fi);                                  // f^(q^6 - 1)
  Fq12 acc; f12_one(acc);
  bool started = false;
  for (int bit = 64 * 32 - 1; bit >= 0; bit--) {        // ^( (q^6 + 1) / r )
    if (started) { Fq12 s; f12_mul(s, acc, acc); acc = s; }
    if ((h_FQ12_HARD_EXP[bit >> 5] >> (bit & 31)) & 1) {
      if (started) { Fq12 s; f12_mul(s, acc, g); acc = s; } else { acc = g; started = true; }
    }
  }
  return f12_is_one(acc);
}

void g2_compress_xyzz(uint8_t* out, const G2Xyzz& p) { G2Affine a; to_affine(a, p); compress(out, a); }

}  // namespace

extern "C" {

int32_t kzg_g2_check(const uint8_t p[96]) {
  G2Affine a;
  return uncompress(a, p) ? KZG_E_BAD_POINT : KZG_OK;
}
void kzg_g2_from_scalar(const uint8_t k_le[32], uint8_t out[96]) {
  G2Affine g; g2_generator(g);
  Fr k; fr_load_le(k, k_le); fr_canon_reduce(k);
  G2Xyzz acc; scalar_mul(acc, g, k.v, 256);
  g2_compress_xyzz(out, acc);
}
int32_t kzg_g2_add(const uint8_t a[96], const uint8_t b[96], uint8_t out[96]) {
  G2Affine pa, pb;
  if (uncompress(pa, a) || uncompress(pb, b)) return KZG_E_BAD_POINT;
  G2Xyzz acc; from_affine(acc, pa); madd(acc, pb, false);
  g2_compress_xyzz(out, acc);
  return KZG_OK;
}
int32_t kzg_g2_sub(const uint8_t a[96], const uint8_t b[96], uint8_t out[96]) {
  G2Affine pa, pb;
  if (uncompress(pa, a) || uncompress(pb, b)) return KZG_E_BAD_POINT;
  G2Xyzz acc; from_affine(acc, pa); madd(acc, pb, true);
  g2_compress_xyzz(out, acc);
  return KZG_OK;
}
int32_t kzg_g2_neg(const uint8_t a[96], uint8_t out[96]) {
  G2Affine pa;
  if (uncompress(pa, a)) return KZG_E_BAD_POINT;
  if (!is_inf(pa)) neg(pa.y, pa.y);
  compress(out, pa);
  return KZG_OK;
}

int32_t kzg_pairing_equal(const uint8_t a1[48], const uint8_t b1[96], const uint8_t a2[48], const uint8_t b2[96],
                          int32_t* equal) {
  G1Affine p1, p2; G2Affine q1, q2;
  if (uncompress(p1, a1) || uncompress(p2, a2) || uncompress(q1, b1) || uncompress(q2, b2)) return KZG_E_BAD_POINT;
  if (!is_inf(p2)) neg(p2.y, p2.y);
  Fq12 f; f12_one(f);
  miller_loop(f, p1, q1);
  miller_loop(f, p2, q2);
  *equal = final_exp_is_one(f) ? 1 : 0;
  return KZG_OK;
}

int32_t kzg_verify_proof(const uint8_t proof[48], const uint8_t commitment[48], const uint8_t z_le[32],
                         const uint8_t y_le[32], const uint8_t g2_tau[96], int32_t* ok) {
  // lhs = e(proof, [s]_2 - [z]_2) ; rhs = e(C - [y]_1, [1]_2)      (src/polynomial.rs:282-291)
  uint8_t zg2[96], rhs2[96], yg1[48], cm[48], one[32] = {1};
  kzg_g2_from_scalar(z_le, zg2);
  int32_t st = kzg_g2_sub(g2_tau, zg2, rhs2);
  if (st != KZG_OK) return st;
  kzg_g1_from_scalar(y_le, yg1);
  st = kzg_g1_sub(commitment, yg1, cm);
  if (st != KZG_OK) return st;
  uint8_t gen2[96];
  kzg_g2_from_scalar(one, gen2);
  return kzg_pairing_equal(proof, rhs2, cm, gen2, ok);
}

}  // extern "C"
