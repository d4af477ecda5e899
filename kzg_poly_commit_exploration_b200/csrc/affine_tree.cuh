// Device helpers shared by the affine pairwise-tree kernels (batch_tree.cu: fixed-base blob batches;
// msm_tree.cu: the first rounds of the bucket accumulation): 48-byte field loads / stores, the denominator
// and the addition of one pair of affine points, complete like blst_p1_add_or_double.
#pragma once
#include "engine.cuh"

namespace kzg {

__device__ __forceinline__ void ld_fq(Fq& o, const Fq* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  o.v[8] = c.x; o.v[9] = c.y; o.v[10] = c.z; o.v[11] = c.w;
}
__device__ __forceinline__ void ld_fq_plain(Fq& o, const Fq* p) {   // data written earlier in the same call chain
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1], c = q[2];
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
  o.v[8] = c.x; o.v[9] = c.y; o.v[10] = c.z; o.v[11] = c.w;
}
__device__ __forceinline__ void st_fq(Fq* p, const Fq& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
  q[2] = make_uint4(v.v[8], v.v[9], v.v[10], v.v[11]);
}

// ---- the pair operation --------------------------------------------------------------------------------
// Denominator of P1 + P2 from lazily loaded coordinates: y is only touched on the rare paths (x = 0 may be the
// infinity marker; x1 = x2 is a doubling or a cancellation).  p == nullptr stands for infinity.
__device__ __forceinline__ void pair_den(Fq& d, const G1Affine* p1, bool n1, const G1Affine* p2, bool n2) {
  set_one(d);
  if (!p1 || !p2) return;
  Fq x1, x2;
  ld_fq_plain(x1, &p1->x); ld_fq_plain(x2, &p2->x);
  const bool z1 = is_zero(x1), z2 = is_zero(x2);
  if (z1 || z2) {
    Fq y1, y2;
    ld_fq_plain(y1, &p1->y); ld_fq_plain(y2, &p2->y);
    if ((z1 && is_zero(y1)) || (z2 && is_zero(y2))) return;
  }
  Fq t; sub(t, x2, x1);
  if (!is_zero(t)) { d = t; return; }
  Fq y1, y2;
  ld_fq_plain(y1, &p1->y); ld_fq_plain(y2, &p2->y);
  if (n1) neg(y1, y1);
  if (n2) neg(y2, y2);
  if (eq(y1, y2) && !is_zero(y1)) dbl(d, y1);
}

// P1 + P2 given dinv = 1 / pair_den(P1, P2); the case analysis mirrors pair_den exactly.
__device__ __forceinline__ void pair_add(G1Affine& o, const G1Affine& a, const G1Affine& b, const Fq& dinv) {
  if (is_inf(a)) { o = b; return; }
  if (is_inf(b)) { o = a; return; }
  Fq lam, t;
  sub(t, b.x, a.x);
  if (is_zero(t)) {
    if (!(eq(a.y, b.y) && !is_zero(a.y))) { set_inf(o); return; }
    sqr(t, a.x);
    dbl(lam, t); add(t, lam, t);          // 3 x^2
    mul(lam, t, dinv);                    // / (2 y)
  } else {
    sub(t, b.y, a.y);
    mul(lam, t, dinv);
  }
  Fq x3;
  sqr(x3, lam); sub(x3, x3, a.x); sub(x3, x3, b.x);
  sub(t, a.x, x3); mul(t, lam, t); sub(o.y, t, a.y);
  o.x = x3;
}

__device__ __forceinline__ void ld_point(G1Affine& o, const G1Affine* p, bool negate) {
  if (!p) { set_inf(o); return; }
  ld_fq_plain(o.x, &p->x); ld_fq_plain(o.y, &p->y);
  if (negate && !is_inf(o)) neg(o.y, o.y);
}

}  // namespace kzg
