// Device code of the affine halving rounds (msm_tree.cu): the pair kernels live in a header so that the round-1 addition
// kernel can ALSO be compiled in a translation unit of its own with the field product out of line (msm_tree_r1.cu).
#pragma once
#include "affine_tree.cuh"

namespace kzg {

constexpr int ST_K = 16;         // consecutive output points per thread

// bucket of output position p: the largest b < M with seg_out[b] <= p (empty buckets share their bound with the
// next non-empty one, and "largest" picks that one)
__device__ __forceinline__ uint32_t bucket_of(const uint32_t* __restrict__ seg_out, uint32_t M, uint32_t p) {
  uint32_t lo = 0, hi = M - 1;
  while (lo < hi) {
    uint32_t mid = (lo + hi + 1) >> 1;
    if (__ldg(seg_out + mid) <= p) lo = mid; else hi = mid - 1;
  }
  return lo;
}

struct PairRef { const G1Affine* p1; const G1Affine* p2; bool n1, n2; };

template <bool R1>
__device__ __forceinline__ PairRef pair_ref(const uint32_t* __restrict__ seg_in, const uint32_t* __restrict__ seg_out,
                                            const uint32_t* __restrict__ vals, const G1Affine* __restrict__ src,
                                            uint32_t b, uint32_t p, uint32_t pitch) {
  const uint32_t q = p - __ldg(seg_out + b);
  const uint32_t s = __ldg(seg_in + b), len = __ldg(seg_in + b + 1) - s;
  const uint32_t e1 = s + 2 * q;
  const bool has2 = 2 * q + 1 < len;
  PairRef r;
  if (R1) {
    const uint32_t v1 = __ldg(vals + e1), v2 = has2 ? __ldg(vals + e1 + 1) : 0u;
    // round 1 gathers by index: the folded-window table (one entry per 128-byte line, MSM_TAB_PITCH), or the dense points
    // of the first level when the bucket sums themselves are reduced as a second bucket problem (pitch 96, msm.cu)
    const char* tb = reinterpret_cast<const char*>(src);
    r.p1 = reinterpret_cast<const G1Affine*>(tb + (size_t)(v1 & 0x7FFFFFFFu) * pitch); r.n1 = (v1 >> 31) != 0;
    r.p2 = has2 ? reinterpret_cast<const G1Affine*>(tb + (size_t)(v2 & 0x7FFFFFFFu) * pitch) : nullptr;
    r.n2 = (v2 >> 31) != 0;
  } else {
    r.p1 = src + e1; r.n1 = false;
    r.p2 = has2 ? src + e1 + 1 : nullptr; r.n2 = false;
  }
  return r;
}

template <bool R1>
__global__ void __launch_bounds__(128)
k_seg_den(const uint32_t* __restrict__ seg_in, const uint32_t* __restrict__ seg_out, uint32_t M,
          const uint32_t* __restrict__ vals, const G1Affine* __restrict__ src, uint64_t T, int K,
          Fq* __restrict__ prefix, Fq* __restrict__ tot, uint32_t pitch) {
  const uint32_t NP = __ldg(seg_out + M);
  // grid-stride over the threads' worth of work: the usual launch has one block per 128 of them, the side-stream launch
  // of round 1 (option den_side) a block or two per SM that stay resident beside another call's addition kernel
#pragma unroll 1
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < T; t += (uint64_t)gridDim.x * blockDim.x) {
    Fq acc; set_one(acc);
    const uint64_t p0 = t * K;
    if (p0 < NP) {
      uint32_t b = bucket_of(seg_out, M, (uint32_t)p0);
#pragma unroll 1
      for (int k = 0; k < K; k++) {
        const uint32_t p = (uint32_t)p0 + k;
        if (p >= NP) break;
        while (__ldg(seg_out + b + 1) <= p) b++;
        PairRef r = pair_ref<R1>(seg_in, seg_out, vals, src, b, p, pitch);
        st_fq(prefix + (size_t)k * T + t, acc);
        if (r.p2) {                                   // the last entry of an odd run passes through: d = 1, no product
          Fq d;
          pair_den(d, r.p1, r.n1, r.p2, r.n2);
          mul(acc, acc, d);
        }
      }
    }
    st_fq(tot + t, acc);
  }
}

template <bool R1>
__global__ void __launch_bounds__(128)
k_seg_add(const uint32_t* __restrict__ seg_in, const uint32_t* __restrict__ seg_out, uint32_t M,
          const uint32_t* __restrict__ vals, const G1Affine* __restrict__ src, uint64_t T, int K,
          const Fq* __restrict__ prefix, Fq* __restrict__ tot_inv, G1Affine* __restrict__ out, uint32_t pitch) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t NP = __ldg(seg_out + M);
  const uint64_t p0 = t * K;
  if (p0 >= NP) return;
  uint32_t bk[ST_K];
  int kmax = 0;
  {
    uint32_t b = bucket_of(seg_out, M, (uint32_t)p0);
    for (int k = 0; k < ST_K; k++) {
      const uint32_t p = (uint32_t)p0 + k;
      if (k >= K || p >= NP) break;
      while (__ldg(seg_out + b + 1) <= p) b++;
      bk[k] = b; kmax = k + 1;
    }
  }
  Fq inv; ld_fq_plain(inv, tot_inv + t);
#pragma unroll 1
  for (int k = kmax - 1; k >= 0; k--) {
    const uint32_t p = (uint32_t)p0 + k;
    PairRef r = pair_ref<R1>(seg_in, seg_out, vals, src, bk[k], p, pitch);
    if (!r.p2) {                                    // passes through (d = 1 in k_seg_den: the thread inverse is untouched)
      G1Affine a; ld_point(a, r.p1, r.n1);
      st_fq(&out[p].x, a.x); st_fq(&out[p].y, a.y);
      continue;
    }
    Fq d, pf, dinv;
    pair_den(d, r.p1, r.n1, r.p2, r.n2);
    ld_fq_plain(pf, prefix + (size_t)k * T + t);
    mul(dinv, inv, pf);
    mul(inv, inv, d);
    G1Affine a, b2, o;
    ld_point(a, r.p1, r.n1);
    ld_point(b2, r.p2, r.n2);
    pair_add(o, a, b2, dinv);
    st_fq(&out[p].x, o.x); st_fq(&out[p].y, o.y);
  }
}

// ---- the same additions with fewer registers, for more warps per scheduler ----------------------------------------
// k_seg_add keeps both points of a pair in registers (116-126 registers: 4 blocks = 16 warps per SM), and ncu shows the
// multiplier pipe only 73-80 % busy there, the rest being fixed-latency dependency stalls that four warps per scheduler
// cannot cover (profiles/prof_seg_add_*_summary.txt).  Here every coordinate is (re)loaded right where it is consumed —
// the lines sit in L1 — so the loop lives in 80 / 96 registers (6 / 5 blocks per SM, a few words of spill), and the rare
// cases (an infinity operand, a doubling, a cancellation) run out of line.
static __device__ __noinline__ void pair_add_rare(G1Affine* out, const G1Affine* p1, bool n1, const G1Affine* p2, bool n2,
                                           const Fq* dinvp) {
  G1Affine a, b, o;
  Fq dinv; ld_fq_plain(dinv, dinvp);
  ld_point(a, p1, n1);
  ld_point(b, p2, n2);
  pair_add(o, a, b, dinv);
  st_fq(&out->x, o.x); st_fq(&out->y, o.y);
}

// VARIANT only names the instantiation: the translation unit that compiles the product out of line (msm_tree_r1.cu) must
// not share a kernel symbol with the inlined one
template <bool R1, int MINB, int VARIANT = 0>
__global__ void __launch_bounds__(128, MINB)
k_seg_add_lean(const uint32_t* __restrict__ seg_in, const uint32_t* __restrict__ seg_out, uint32_t M,
               const uint32_t* __restrict__ vals, const G1Affine* __restrict__ src, uint64_t T, int K,
               const Fq* __restrict__ prefix, Fq* __restrict__ tot_inv, G1Affine* __restrict__ out, uint32_t pitch) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint32_t NP = __ldg(seg_out + M);
  const uint64_t p0 = t * K;
  if (p0 >= NP) return;
  const int kmax = (int)min((uint64_t)K, (uint64_t)NP - p0);
  Fq inv; ld_fq_plain(inv, tot_inv + t);
  uint32_t b = bucket_of(seg_out, M, (uint32_t)(p0 + kmax - 1));
#pragma unroll 1
  for (int k = kmax - 1; k >= 0; k--) {
    const uint32_t p = (uint32_t)p0 + k;
    while (__ldg(seg_out + b) > p) b--;            // the bucket of p: positions are walked downwards
    const PairRef r = pair_ref<R1>(seg_in, seg_out, vals, src, b, p, pitch);
    if (r.p2 == nullptr) {                          // the last entry of an odd run passes through: a copy, no product
      Fq x, y; ld_fq_plain(x, &r.p1->x); ld_fq_plain(y, &r.p1->y);   // (d = 1 in k_seg_den: the thread inverse is untouched)
      if (R1) { Fq m; neg(m, y); if (r.n1 && !(is_zero(x) && is_zero(y))) y = m; }
      st_fq(&out[p].x, x); st_fq(&out[p].y, y);
      continue;
    }
    Fq d, dinv;
    bool rare;
    {
      Fq x1; ld_fq_plain(x1, &r.p1->x); ld_fq_plain(d, &r.p2->x);
      rare = is_zero(x1) || is_zero(d);            // x = 0 may be the infinity marker
      sub(d, d, x1);
      rare = rare || is_zero(d);                   // doubling or cancellation
    }
    if (rare) pair_den(d, r.p1, r.n1, r.p2, r.n2);  // the same case analysis k_seg_den ran
    { Fq pf; ld_fq_plain(pf, prefix + (size_t)k * T + t); mul(dinv, inv, pf); }
    mul(inv, inv, d);
    if (rare) {
      st_fq(tot_inv + t, dinv);                    // this thread's own slot: read once above, free since
      pair_add_rare(out + p, r.p1, r.n1, r.p2, r.n2, tot_inv + t);
      continue;
    }
    Fq lam;
    {
      Fq y1, y2; ld_fq_plain(y1, &r.p1->y); ld_fq_plain(y2, &r.p2->y);
      if (R1) { Fq m; neg(m, y1); if (r.n1) y1 = m; neg(m, y2); if (r.n2) y2 = m; }
      sub(y2, y2, y1); mul(lam, y2, dinv);
    }
    Fq x3;
    sqr(x3, lam);
    { Fq x; ld_fq_plain(x, &r.p1->x); sub(x3, x3, x); ld_fq_plain(x, &r.p2->x); sub(x3, x3, x); }
    st_fq(&out[p].x, x3);
    {
      Fq x; ld_fq_plain(x, &r.p1->x); sub(x3, x, x3); mul(x3, lam, x3);
      ld_fq_plain(x, &r.p1->y);
      if (R1) { Fq m; neg(m, x); if (r.n1) x = m; }
      sub(x3, x3, x);
    }
    st_fq(&out[p].y, x3);
  }
}

}  // namespace kzg
