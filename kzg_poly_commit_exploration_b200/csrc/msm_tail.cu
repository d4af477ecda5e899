// K2 tail — bucket reduction sum_b (b+1) * B_b and plain point sums (latency-bound part of the MSM).
//
// These kernels run a handful of warps through long chains of dependent XYZZ additions.  With the field
// multiplication inlined an addition is ~84 KB of straight-line code executed ONCE per call, and a lone warp
// then waits on instruction fetch, not on arithmetic (k_reduce_l3's single-thread Horner ran at 3.8 us per Fq
// product against 1.4 us inside a resident loop).  This translation unit therefore compiles the Fq product
// as a real function (KZG_FQ_NOINLINE: arguments in registers, one 6 KB copy), which keeps the whole tail
// inside the instruction cache.
#define KZG_FQ_NOINLINE 1
#include <algorithm>
#include "engine.cuh"

namespace kzg {

// ---- 5. reduce ---------------------------------------------------------------------------------
// sum_i (i+1) * A_i over M buckets, warp-cooperatively.  A warp owns 32 consecutive buckets, one per
// lane: a 5-step suffix scan through shuffles gives R_l = sum_{m>=l} A_m, hence S = R_0 and
// W = sum_l l*A_l = sum_{l>=1} R_l (a 5-step tree).  Then sum_i i*A_i = sum_k W_k + 32 * sum_k k*S_k
// and the second term is the same problem on the 32x smaller array S: log_32(M) levels of ten
// sequential additions each instead of a 2L-long running sum per thread.
__device__ __forceinline__ void shfl_down_pt(G1Xyzz& o, const G1Xyzz& p, int d) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    o.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], d);
    o.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], d);
    o.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], d);
    o.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], d);
  }
}

__global__ void __launch_bounds__(128)
k_reduce_warp(const G1Xyzz* __restrict__ A, uint32_t M, G1Xyzz* __restrict__ W_out, G1Xyzz* __restrict__ S_out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t k = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t chunks = (M + 31) >> 5;
  if (k >= chunks) return;                       // whole warps exit together
  const uint32_t idx = k * 32 + lane;
  G1Xyzz R;
  if (idx < M) R = A[idx]; else set_inf(R);
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {             // suffix scan
    G1Xyzz o; shfl_down_pt(o, R, d);
    if (lane + d < 32) add(R, o);
  }
  if (lane == 0) S_out[k] = R;
  G1Xyzz X = R;
  if (lane == 0) set_inf(X);
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {            // tree sum of R_1..R_31
    G1Xyzz o; shfl_down_pt(o, X, d);
    if (lane < d) add(X, o);
  }
  if (lane == 0) W_out[k] = X;
}

// out[k] = sum of in[32k .. 32k+32), one warp per output
__global__ void __launch_bounds__(128)
k_sum_warp(const G1Xyzz* __restrict__ in, uint64_t count, G1Xyzz* __restrict__ out) {
  const uint32_t lane = threadIdx.x & 31;
  const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t chunks = (count + 31) >> 5;
  if (k >= chunks) return;
  const uint64_t idx = k * 32 + lane;
  G1Xyzz X;
  if (idx < count) X = in[idx]; else set_inf(X);
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o; shfl_down_pt(o, X, d);
    if (lane < d) add(X, o);
  }
  if (lane == 0) out[k] = X;
}

// result = S_total + Wsum_0 + 32*(Wsum_1 + 32*(Wsum_2 + ...)) : single thread, a handful of operations
__global__ void k_reduce_final(const G1Xyzz* __restrict__ lvl, int nlevels, const G1Xyzz* __restrict__ s_total,
                               G1Xyzz* __restrict__ out) {
  G1Xyzz acc = lvl[nlevels - 1];
  for (int l = nlevels - 2; l >= 0; l--) {
    for (int k = 0; k < 5; k++) dbl(acc);
    G1Xyzz w = lvl[l]; add(acc, w);
  }
  G1Xyzz st = s_total[0]; add(acc, st);
  out[0] = acc;
}

// ---- fused tail (M <= 32768 buckets): three launches, ~47 dependent additions instead of ~71 -----------
// The tail is latency-bound (a dependent XYZZ addition costs a lone warp ~15-20 us), so what matters is
// the LENGTH of the dependency chain, not the work: independent chains run in different warps.
__device__ __forceinline__ void warp_suffix_scan(G1Xyzz& R, uint32_t lane) {
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1Xyzz o; shfl_down_pt(o, R, d);
    if (lane + d < 32) add(R, o);
  }
}
__device__ __forceinline__ void warp_tree_sum(G1Xyzz& X, uint32_t lane) {
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o; shfl_down_pt(o, X, d);
    if (lane < d) add(X, o);
  }
}

// level 1: one warp owns 32 consecutive buckets, one per lane.  A lane first folds the T partials of its
// bucket serially (every lane does useful work; a shuffle tree here would idle most lanes and, at 8x the
// threads, no longer fit in one wave — measured 1.1 ms slower), then the warp turns the 32 bucket sums into
// S_k = sum_l A_l and W_k = sum_l l*A_l (suffix scan + tree).  Blocks are single warps so that all
// M/32 chains are resident at once: the tail is latency-bound.
__global__ void __launch_bounds__(32)
k_reduce_l1(const G1Xyzz* __restrict__ parts, uint32_t M, uint32_t T, G1Xyzz* __restrict__ W_out,
            G1Xyzz* __restrict__ S_out, uint32_t pool_stride) {
  // blockIdx.y = bucket set (reduce_buckets_multi): its partials follow the previous set's, its outputs live pool_stride on
  parts += (size_t)blockIdx.y * M * T; W_out += (size_t)blockIdx.y * pool_stride; S_out += (size_t)blockIdx.y * pool_stride;
  const uint32_t lane = threadIdx.x;
  const uint32_t b = blockIdx.x * 32 + lane;
  G1Xyzz X; set_inf(X);
  if (b < M) {
    X = parts[(size_t)b * T];
#pragma unroll 1
    for (uint32_t k = 1; k < T; k++) { G1Xyzz o = parts[(size_t)b * T + k]; add(X, o); }
  }
  warp_suffix_scan(X, lane);
  if (lane == 0) S_out[blockIdx.x] = X;
  if (lane == 0) set_inf(X);
  warp_tree_sum(X, lane);
  if (lane == 0) W_out[blockIdx.x] = X;
}

// level 1 for calls that run for THROUGHPUT (several in flight, or fused): with the machine shared between calls what
// counts is the work, not the length of the chain.  k_reduce_l1 spends 10 warp-wide additions on every 32 buckets (the
// suffix scan and the tree keep half the lanes idle); here a lane first walks 2^LK consecutive buckets with the
// running-sum recurrence (acc += B_i; run += acc: two additions per bucket), and the warp-level steps are paid once per
// 32 * 2^LK buckets: 2 + 11 / 2^LK additions (+ LK doublings per lane) per bucket instead of 10.
//   lane: S_l = sum_i B_{b0+i}, W_l = sum_i i * B_{b0+i};  warp: S = sum_l S_l, W = 2^LK * sum_l l * S_l + sum_l W_l
template <int LK>
__global__ void __launch_bounds__(32)
k_reduce_l1k(const G1Xyzz* __restrict__ parts, uint32_t M, uint32_t T, G1Xyzz* __restrict__ W_out,
             G1Xyzz* __restrict__ S_out, uint32_t pool_stride) {
  parts += (size_t)blockIdx.y * M * T; W_out += (size_t)blockIdx.y * pool_stride; S_out += (size_t)blockIdx.y * pool_stride;
  constexpr uint32_t K = 1u << LK;
  const uint32_t lane = threadIdx.x;
  const uint32_t b0 = (blockIdx.x * 32 + lane) * K;
  G1Xyzz acc, run; set_inf(acc); set_inf(run);
#pragma unroll 1
  for (int i = (int)K - 1; i >= 0; i--) {
    const uint32_t b = b0 + (uint32_t)i;
    if (b < M) {
#pragma unroll 1
      for (uint32_t k = 0; k < T; k++) { G1Xyzz o = parts[(size_t)b * T + k]; add(acc, o); }
    }
    if (i > 0) add(run, acc);
  }
  warp_suffix_scan(acc, lane);                // R_l = sum_{m >= l} S_m
  if (lane == 0) S_out[blockIdx.x] = acc;
  // sum_l (2^LK * l * S_l + W_l) = sum_l Y_l with Y_l = 2^LK * R_l (l >= 1) + W_l: ONE tree instead of two
  if (lane == 0) set_inf(acc);
#pragma unroll 1
  for (int k = 0; k < LK; k++) dbl(acc);
  add(acc, run);
  warp_tree_sum(acc, lane);
  if (lane == 0) W_out[blockIdx.x] = acc;
}

// level 2: chunk k of 32 level-1 outputs.  Even warps run the S chain (-> S1_k, W1_k), odd warps fold the
// 32 W0 values of the same chunk (-> V_k), so the two chains overlap instead of queueing in one warp.
__global__ void __launch_bounds__(128)
k_reduce_l2(const G1Xyzz* __restrict__ S0, const G1Xyzz* __restrict__ W0, uint32_t m,
            G1Xyzz* __restrict__ S1, G1Xyzz* __restrict__ W1, G1Xyzz* __restrict__ V) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t wid = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t k = wid >> 1;
  const uint32_t chunks = (m + 31) >> 5;
  if (k >= chunks) return;
  const uint32_t idx = k * 32 + lane;
  G1Xyzz X;
  if (wid & 1) {
    if (idx < m) X = W0[idx]; else set_inf(X);
    warp_tree_sum(X, lane);
    if (lane == 0) V[k] = X;
  } else {
    if (idx < m) X = S0[idx]; else set_inf(X);
    warp_suffix_scan(X, lane);
    if (lane == 0) S1[k] = X;
    if (lane == 0) set_inf(X);
    warp_tree_sum(X, lane);
    if (lane == 0) W1[k] = X;
  }
}

// level 3 + Horner: m2 <= 32.  warp 0: S1 -> (S_total, W2); warp 1: sum W1; warp 2: sum V (= sum W0);
// thread 0: out = S_total + sumW0 + 32*(sumW1 + 32*(W2 ...)), then the shallower pre-summed levels.
__global__ void __launch_bounds__(96)
k_reduce_l3(const G1Xyzz* __restrict__ S1, const G1Xyzz* __restrict__ W1, const G1Xyzz* __restrict__ V, uint32_t m2,
            const G1Xyzz* __restrict__ pre, int npre, G1Xyzz* __restrict__ out) {
  __shared__ G1Xyzz sh[4];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  G1Xyzz X;
  const G1Xyzz* src = w == 0 ? S1 : (w == 1 ? W1 : V);
  if (lane < m2) X = src[lane]; else set_inf(X);
  if (w == 0) {
    warp_suffix_scan(X, lane);
    if (lane == 0) sh[0] = X;          // S_total
    if (lane == 0) set_inf(X);
    warp_tree_sum(X, lane);            // W2
  } else {
    warp_tree_sum(X, lane);
    if (lane == 0) sh[w] = X;          // sh[1] = sum W1, sh[2] = sum W0
  }
  __syncthreads();
  if (threadIdx.x != 0) return;
  G1Xyzz acc = X;                      // W2
  for (int l = 1; l >= 0; l--) {
    for (int k = 0; k < 5; k++) dbl(acc);
    G1Xyzz t = sh[l == 1 ? 1 : 2]; add(acc, t);
  }
  for (int l = npre - 1; l >= 0; l--) {
    for (int k = 0; k < 5; k++) dbl(acc);
    G1Xyzz t = pre[l]; add(acc, t);
  }
  G1Xyzz st = sh[0]; add(acc, st);
  out[0] = acc;
}

// out[k] = sum of in[k L .. k L + L)
__global__ void __launch_bounds__(128)
k_sum_chunks(const G1Xyzz* __restrict__ in, uint64_t count, uint32_t L, G1Xyzz* __restrict__ out) {
  uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t chunks = (count + L - 1) / L;
  if (k >= chunks) return;
  G1Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (uint32_t l = 0; l < L; l++) {
    uint64_t idx = k * L + l;
    if (idx < count) { G1Xyzz b = in[idx]; add(acc, b); }
  }
  out[k] = acc;
}

__global__ void k_set_inf_tail(G1Xyzz* out, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { G1Xyzz p; set_inf(p); out[i] = p; }
}

// ---- 4-warp cooperative tail ---------------------------------------------------------------------------
// A dependent XYZZ addition is 14 Fq products in sequence, but only 4 deep: u1,u2,s1,s2 | pp,rr,zz,zzz |
// ppp,q,zz3 | t,s1ppp,zzz3.  In the tail nothing else can fill the machine, so the four warps of a block hold
// IDENTICAL copies of one point per lane and each computes one product of the current stage; results travel
// through 6 KB of shared memory.  An addition then costs 4 product latencies instead of 14, a doubling 3
// instead of 9.  P = Q / P = -Q inside an addition (rare in the tail) fall back to the plain formula, decided
// block-wide so that every thread meets the same barriers.
struct CoopBuf { Fq v[4][32]; };

__device__ __forceinline__ void coop_stage(Fq& r0, Fq& r1, Fq& r2, Fq& r3, const Fq& a0, const Fq& b0, const Fq& a1,
                                           const Fq& b1, const Fq& a2, const Fq& b2, const Fq& a3, const Fq& b3,
                                           CoopBuf* sh, uint32_t w, uint32_t lane) {
  Fq x, y;
  if (w == 0) { x = a0; y = b0; } else if (w == 1) { x = a1; y = b1; } else if (w == 2) { x = a2; y = b2; } else { x = a3; y = b3; }
  Fq pr; mul(pr, x, y);
  sh->v[w][lane] = pr;
  __syncthreads();
  r0 = sh->v[0][lane]; r1 = sh->v[1][lane]; r2 = sh->v[2][lane]; r3 = sh->v[3][lane];
  __syncthreads();
}

// a += b (every warp of the block calls this with the same per-lane operands)
__device__ __noinline__ void add4(G1Xyzz& a, const G1Xyzz& b, bool active, CoopBuf* sh, uint32_t w, uint32_t lane) {
  const bool inf_a = is_inf(a), inf_b = is_inf(b);
  Fq u1, u2, s1, s2, p, r, pp, rr, zz12, zzz12, ppp, q, zz3, zzz3, t, s1p, dummy;
  coop_stage(u1, u2, s1, s2, a.x, b.zz, b.x, a.zz, a.y, b.zzz, b.y, a.zzz, sh, w, lane);
  sub(p, u2, u1);
  sub(r, s2, s1);
  // lanes whose result is discarded (shuffles past the end of the warp hand a lane its own value) must not
  // drag the block into the slow path
  const bool special = active && !inf_a && !inf_b && is_zero(p);
  coop_stage(pp, rr, zz12, zzz12, p, p, r, r, a.zz, b.zz, a.zzz, b.zzz, sh, w, lane);
  coop_stage(ppp, q, zz3, dummy, p, pp, u1, pp, zz12, pp, p, p, sh, w, lane);
  Fq x3; sub(x3, rr, ppp); sub(x3, x3, q); sub(x3, x3, q);
  Fq qx; sub(qx, q, x3);
  coop_stage(t, s1p, zzz3, dummy, r, qx, s1, ppp, zzz12, ppp, p, p, sh, w, lane);
  G1Xyzz res;
  res.x = x3; sub(res.y, t, s1p); res.zz = zz3; res.zzz = zzz3;
  if (__syncthreads_or(special ? 1 : 0)) {
    G1Xyzz slow = a; add(slow, b);          // plain complete formula (doubling / cancellation)
    if (special) res = slow;
  }
  if (inf_b) res = a;
  else if (inf_a) res = b;
  a = res;
}

// a = 2a
__device__ __noinline__ void dbl4(G1Xyzz& a, CoopBuf* sh, uint32_t w, uint32_t lane) {
  const bool inf_a = is_inf(a);
  Fq u, v, xx, m, wv, sv, mm, zzn, t, wy, zzzn, dummy;
  dbl(u, a.y);
  coop_stage(v, xx, dummy, dummy, u, u, a.x, a.x, u, u, u, u, sh, w, lane);
  dbl(m, xx); add(m, m, xx);
  coop_stage(wv, sv, mm, zzn, u, v, a.x, v, m, m, v, a.zz, sh, w, lane);
  Fq x3; sub(x3, mm, sv); sub(x3, x3, sv);
  Fq sx; sub(sx, sv, x3);
  coop_stage(t, wy, zzzn, dummy, m, sx, wv, a.y, wv, a.zzz, u, u, sh, w, lane);
  if (!inf_a) { a.x = x3; sub(a.y, t, wy); a.zz = zzn; a.zzz = zzzn; }
}

__device__ __forceinline__ void coop_suffix_scan(G1Xyzz& R, CoopBuf* sh, uint32_t w, uint32_t lane) {
#pragma unroll 1
  for (int d = 1; d < 32; d <<= 1) {
    G1Xyzz o; shfl_down_pt(o, R, d);
    G1Xyzz n = R; add4(n, o, lane + d < 32, sh, w, lane);
    if (lane + d < 32) R = n;
  }
}
__device__ __forceinline__ void coop_tree_sum(G1Xyzz& X, CoopBuf* sh, uint32_t w, uint32_t lane) {
#pragma unroll 1
  for (int d = 16; d >= 1; d >>= 1) {
    G1Xyzz o; shfl_down_pt(o, X, d);
    G1Xyzz n = X; add4(n, o, lane < d, sh, w, lane);
    if (lane < d) X = n;
  }
}

// level 1, cooperative: a block (4 warps) owns 32 consecutive buckets
__global__ void __launch_bounds__(128)
k_reduce_l1c(const G1Xyzz* __restrict__ parts, uint32_t M, uint32_t T, G1Xyzz* __restrict__ W_out,
             G1Xyzz* __restrict__ S_out) {
  __shared__ CoopBuf sh;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t b = blockIdx.x * 32 + lane;
  G1Xyzz X; set_inf(X);
  if (b < M) X = parts[(size_t)b * T];
#pragma unroll 1
  for (uint32_t k = 1; k < T; k++) {
    G1Xyzz o; set_inf(o);
    if (b < M) o = parts[(size_t)b * T + k];
    add4(X, o, true, &sh, w, lane);
  }
  coop_suffix_scan(X, &sh, w, lane);
  if (threadIdx.x == 0) S_out[blockIdx.x] = X;
  if (lane == 0) set_inf(X);
  coop_tree_sum(X, &sh, w, lane);
  if (threadIdx.x == 0) W_out[blockIdx.x] = X;
}

// level 2, cooperative: block 2k runs the S chain of chunk k, block 2k+1 folds its W0 values
__global__ void __launch_bounds__(128)
k_reduce_l2c(const G1Xyzz* __restrict__ S0, const G1Xyzz* __restrict__ W0, uint32_t m,
             G1Xyzz* __restrict__ S1, G1Xyzz* __restrict__ W1, G1Xyzz* __restrict__ V, uint32_t pool_stride) {
  __shared__ CoopBuf sh;
  { const size_t o = (size_t)blockIdx.y * pool_stride; S0 += o; W0 += o; S1 += o; W1 += o; V += o; }
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t k = blockIdx.x >> 1;
  const uint32_t idx = k * 32 + lane;
  G1Xyzz X;
  if (blockIdx.x & 1) {
    if (idx < m) X = W0[idx]; else set_inf(X);
    coop_tree_sum(X, &sh, w, lane);
    if (threadIdx.x == 0) V[k] = X;
  } else {
    if (idx < m) X = S0[idx]; else set_inf(X);
    coop_suffix_scan(X, &sh, w, lane);
    if (threadIdx.x == 0) S1[k] = X;
    if (lane == 0) set_inf(X);
    coop_tree_sum(X, &sh, w, lane);
    if (threadIdx.x == 0) W1[k] = X;
  }
}

// level 3, cooperative, three blocks: 0: S1 -> (S_total, W2); 1: sum W1; 2: sum V.  res[0..3] = S_total, W2, sumW1, sumV
__global__ void __launch_bounds__(128)
k_reduce_l3c(const G1Xyzz* __restrict__ S1, const G1Xyzz* __restrict__ W1, const G1Xyzz* __restrict__ V, uint32_t m2,
             G1Xyzz* __restrict__ res, uint32_t pool_stride) {
  __shared__ CoopBuf sh;
  { const size_t o = (size_t)blockIdx.y * pool_stride; S1 += o; W1 += o; V += o; res += o; }
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const G1Xyzz* src = blockIdx.x == 0 ? S1 : (blockIdx.x == 1 ? W1 : V);
  G1Xyzz X;
  if (lane < m2) X = src[lane]; else set_inf(X);
  if (blockIdx.x == 0) {
    coop_suffix_scan(X, &sh, w, lane);
    if (threadIdx.x == 0) res[0] = X;
    if (lane == 0) set_inf(X);
    coop_tree_sum(X, &sh, w, lane);
    if (threadIdx.x == 0) res[1] = X;
  } else {
    coop_tree_sum(X, &sh, w, lane);
    if (threadIdx.x == 0) res[blockIdx.x + 1] = X;
  }
}

// Horner, cooperative: out = S_total + sumW0 + 32*(sumW1 + 32*(W2)), then the shallower pre-summed levels
__global__ void __launch_bounds__(128)
k_reduce_horner(const G1Xyzz* __restrict__ res, const G1Xyzz* __restrict__ pre, int npre, G1Xyzz* __restrict__ out,
                uint32_t pool_stride, int l1_shift) {
  // l1_shift = log2 of the buckets per level-1 chunk (5, or 5 + LK after k_reduce_l1k): the weight of the step that adds sum W0
  __shared__ CoopBuf sh;
  res += (size_t)blockIdx.y * pool_stride; out += blockIdx.y;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  G1Xyzz acc = res[1];
#pragma unroll 1
  for (int l = 0; l < 2 + npre; l++) {
    const int shift = l == 1 ? l1_shift : 5;
#pragma unroll 1
    for (int k = 0; k < shift; k++) dbl4(acc, &sh, w, lane);
    G1Xyzz t = l == 0 ? res[2] : (l == 1 ? res[3] : pre[npre - 1 - (l - 2)]);
    add4(acc, t, true, &sh, w, lane);
  }
  G1Xyzz st = res[0];
  add4(acc, st, true, &sh, w, lane);
  if (threadIdx.x == 0) out[0] = acc;
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// out[k] = sum of in[32k .. 32k+32), one cooperative block per output (few, latency-bound sums: the multi-GPU combine)
__global__ void __launch_bounds__(128)
k_sum_coop(const G1Xyzz* __restrict__ in, uint64_t count, G1Xyzz* __restrict__ out) {
  __shared__ CoopBuf sh;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t idx = (uint64_t)blockIdx.x * 32 + lane;
  G1Xyzz X;
  if (idx < count) X = in[idx]; else set_inf(X);
  coop_tree_sum(X, &sh, w, lane);
  if (threadIdx.x == 0) out[blockIdx.x] = X;
}

// ---- multi-GPU: exchange + fold + compress in ONE kernel over NVLink peer memory ---------------------------------
// Point-range sharding leaves one 192-byte XYZZ partial sum per GPU.  Instead of an NCCL all-gather followed by a
// fold kernel and a compression kernel, every rank runs this single block: lane r of warp 0 STORES the rank's partial
// straight into rank r's mailbox (peer memory mapped through CUDA IPC, NVLink 5 / NVSwitch), publishes it with a
// release store of the call's sequence number, then spins (acquire loads) on its own mailbox until all `world`
// partials of this call have landed; the four warps fold them with cooperative additions and thread 0 converts to
// affine and compresses.  Mailbox slots are KZG_P2P_DEPTH = 2 F deep (F = KZG_MAX_INFLIGHT = lanes x fused calls: the
// tickets a rank can have out), indexed by the call number: a rank submits k + F only after waiting for k, k completes on
// a rank only when EVERY rank has published k, and a rank publishes k only after it finished k - F; so when call k + F
// overwrites the slot of call k - F every rank is done reading it.  The collective calls of a fused group are the blocks
// of ONE launch (block b = call seq + b, its own slot).  A bounded spin (20 s by default, option p2p_timeout_ms, counted with the device's
// own clock rate) turns a dead peer into an error instead of a hang, and that error is COLLECTIVE: the rank that gives
// up writes a sticky word into every peer's mailbox, peers check it while they wait and once more before they return,
// and the contexts refuse further collective calls (their sequence numbers are no longer in step).
__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__global__ void __launch_bounds__(128)
k_exchange_combine(const G1Xyzz* __restrict__ mine, PeerTable peers, int world, int rank, uint32_t seq,
                   long long timeout_cycles, uint8_t* __restrict__ out_c48, uint32_t* __restrict__ status, uint32_t out_stride) {
  __shared__ CoopBuf sh;
  // block b = collective call number seq + b (calls fused into one bucket problem finish together: one launch for all)
  mine += blockIdx.x; seq += blockIdx.x;
  out_c48 += (size_t)blockIdx.x * out_stride;
  status = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(status) + (size_t)blockIdx.x * out_stride);
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t par = seq & (KZG_P2P_DEPTH - 1);
  PeerMailbox* my = peers.box[rank];
  const bool talker = w == 0 && lane < (uint32_t)world;
  int timed_out = 0, peer_dead = 0;
  if (talker) {
    const uint4* src = reinterpret_cast<const uint4*>(mine);
    PeerMailbox* mb = peers.box[lane];
    volatile uint4* dst = reinterpret_cast<volatile uint4*>(&mb->slot[par][rank]);
#pragma unroll
    for (int i = 0; i < 12; i++) { uint4 v = src[i]; dst[i].x = v.x; dst[i].y = v.y; dst[i].z = v.z; dst[i].w = v.w; }
    __threadfence_system();
    st_release_sys(&mb->flag[par][rank], seq);
    const long long t0 = clock64();
    while ((int32_t)(ld_acquire_sys(&my->flag[par][lane]) - seq) < 0) {
      if (ld_acquire_sys(&my->dead[lane]) != 0) { peer_dead = 1; break; }      // that rank has given up: do not wait for it
      if (clock64() - t0 > timeout_cycles) { timed_out = 1; break; }
      __nanosleep(64);
    }
    // everything has landed — but a peer may have stopped waiting for US before we published: then it returned an
    // error for this call, and so must we (all ranks agree on the outcome of a collective call)
    if (!timed_out && !peer_dead && ld_acquire_sys(&my->dead[lane]) != 0) peer_dead = 1;
  }
  const int any_timeout = __syncthreads_or(timed_out);
  const int any_dead = __syncthreads_or(peer_dead);
  if (any_timeout || any_dead) {
    // give the call up on EVERY rank: the sticky word tells peers that are still waiting (or arrive later) to fail too.
    // What remains is the window of one NVLink round trip between a rank's last look at `dead` and a peer's store to it.
    if (talker) st_release_sys(&peers.box[lane]->dead[rank], seq ? seq : 1u);
    if (threadIdx.x == 0) *status = any_timeout ? 1u : 2u;
    return;
  }
  G1Xyzz X;
  if (lane < (uint32_t)world) {
    const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(&my->slot[par][lane]);
#pragma unroll
    for (int i = 0; i < 12; i++) { X.x.v[i] = q[i]; X.y.v[i] = q[12 + i]; X.zz.v[i] = q[24 + i]; X.zzz.v[i] = q[36 + i]; }
  } else set_inf(X);
  coop_tree_sum(X, &sh, w, lane);
  if (threadIdx.x == 0) {
    G1Affine a; to_affine(a, X);
    uint8_t buf[48];
    compress(buf, a);
    for (int k = 0; k < 48; k++) out_c48[k] = buf[k];
    *status = 0u;
  }
}

int32_t exchange_combine(kzg_ctx* ctx, const kzg_ctx* owner, uint32_t seq, const G1Xyzz* mine_dev, uint8_t* out_c48_dev,
                         uint32_t* status_dev, int calls, uint32_t out_stride) {
  if (owner->p2p_world < 1 || !owner->p2p_connected) return fail(ctx, KZG_E_BAD_ARG, "peer mailboxes not connected (kzg_p2p_init / kzg_p2p_connect)");
  PeerTable t;
  for (int r = 0; r < KZG_MAX_PEERS; r++) t.box[r] = r < owner->p2p_world ? reinterpret_cast<PeerMailbox*>(owner->p2p_box[r]) : nullptr;
  k_exchange_combine<<<(unsigned)calls, 128, 0, ctx->stream>>>(mine_dev, t, owner->p2p_world, owner->p2p_rank, seq,
                                                               owner->p2p_timeout_cycles, out_c48_dev, status_dev, out_stride);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

int32_t sum_points(kzg_ctx* ctx, const G1Xyzz* in_dev, uint64_t count, G1Xyzz* out_dev) {
  if (count == 0) {
    k_set_inf_tail<<<1, 32, 0, ctx->stream>>>(out_dev, 1);
    KZG_LAUNCHED(ctx);
    return KZG_OK;
  }
  void *a, *b;
  uint64_t c1 = (count + 31) / 32;
  KZG_TRY(scratch_get(ctx, SCR_RED_A, (c1 + 1) * sizeof(G1Xyzz), &a));
  KZG_TRY(scratch_get(ctx, SCR_RED_B, ((c1 + 31) / 32 + 1) * sizeof(G1Xyzz), &b));
  const G1Xyzz* src = in_dev;
  G1Xyzz* bufs[2] = {reinterpret_cast<G1Xyzz*>(a), reinterpret_cast<G1Xyzz*>(b)};
  int which = 0;
  while (true) {
    uint64_t chunks = (count + 31) / 32;
    G1Xyzz* dst = chunks == 1 ? out_dev : bufs[which];
    if (chunks <= 64 && ctx->opt_coop_tail) k_sum_coop<<<(unsigned)chunks, 128, 0, ctx->stream>>>(src, count, dst);
    else k_sum_warp<<<blocks_for(chunks * 32, 128), 128, 0, ctx->stream>>>(src, count, dst);
    KZG_LAUNCHED(ctx);
    if (chunks == 1) break;
    src = dst; count = chunks; which ^= 1;
  }
  return KZG_OK;
}

// buckets a lane of level 1 walks serially (log2): 0 = k_reduce_l1 (shortest chain, a call on its own), else k_reduce_l1k
static int l1_chunk_log(const kzg_ctx* ctx, uint32_t M) {
  if (ctx->opt_reduce_lk >= 0) return std::min(3, ctx->opt_reduce_lk);
  if (!ctx->throughput_mode) return M > 32768 ? 1 : 0;      // a call on its own: the shortest chain that still fits the fused tail
  return M > 32768 ? 3 : (M >= 8192 ? 2 : (M >= 2048 ? 1 : 0));
}
static void launch_l1(kzg_ctx* ctx, int lk, dim3 grid, const G1Xyzz* parts, uint32_t M, uint32_t T, G1Xyzz* W0, G1Xyzz* S0,
                      uint32_t stride) {
  if (lk == 1) k_reduce_l1k<1><<<grid, 32, 0, ctx->stream>>>(parts, M, T, W0, S0, stride);
  else if (lk == 2) k_reduce_l1k<2><<<grid, 32, 0, ctx->stream>>>(parts, M, T, W0, S0, stride);
  else if (lk == 3) k_reduce_l1k<3><<<grid, 32, 0, ctx->stream>>>(parts, M, T, W0, S0, stride);
  else k_reduce_l1<<<grid, 32, 0, ctx->stream>>>(parts, M, T, W0, S0, stride);
}

// sum_b (b+1) * A_b -> result_dev, where A_b = sum_k parts[b*T + k], T = 2^logT partials per bucket
int32_t reduce_buckets(kzg_ctx* ctx, const G1Xyzz* parts, uint32_t M, uint32_t logT, G1Xyzz* result_dev) {
  // shallow levels (only for M > 32768, i.e. windows wider than 16 bits) run one launch per step and leave
  // their W sums in pre[]; the last three levels are the fused tail.
  void* ws;
  uint32_t shallow = 0;
  // the fused tail takes up to 1024 level-1 chunks: 32768 buckets with 32-bucket chunks, more with k_reduce_l1k's wider ones
  const uint32_t fused_max = ctx->opt_coop_tail == 1 ? (32768u << l1_chunk_log(ctx, M)) : 32768u;
  for (uint32_t m = M; m > fused_max; m = (m + 31) / 32) shallow++;
  size_t cap = (size_t)(M + 31) / 32 * 2 * 2 + 4096 + 16;
  KZG_TRY(scratch_get(ctx, SCR_XYZZ_TMP, cap * sizeof(G1Xyzz), &ws));
  G1Xyzz* pool = reinterpret_cast<G1Xyzz*>(ws);
  G1Xyzz* pre = pool; pool += 8;
  const G1Xyzz* cur = parts;
  uint32_t m = M;
  if (shallow) {
    if (logT) {          // fold the T partials first: the generic level kernel takes one value per bucket
      void* bk;
      KZG_TRY(scratch_get(ctx, SCR_BUCKETS, (size_t)M * sizeof(G1Xyzz), &bk));
      k_sum_chunks<<<blocks_for(M, 128), 128, 0, ctx->stream>>>(parts, (uint64_t)M << logT, 1u << logT,
                                                                reinterpret_cast<G1Xyzz*>(bk));
      KZG_LAUNCHED(ctx);
      cur = reinterpret_cast<G1Xyzz*>(bk); logT = 0;
    }
    for (uint32_t l = 0; l < shallow; l++) {
      uint32_t k = (m + 31) / 32;
      G1Xyzz* W = pool; pool += k;
      G1Xyzz* S = pool; pool += k;
      k_reduce_warp<<<blocks_for((uint64_t)k * 32, 128), 128, 0, ctx->stream>>>(cur, m, W, S);
      KZG_LAUNCHED(ctx);
      KZG_TRY(sum_points(ctx, W, k, pre + l));
      cur = S; m = k;
    }
  }
  const int lk = (shallow == 0 && ctx->opt_coop_tail == 1) ? l1_chunk_log(ctx, m) : 0;
  const uint32_t m1 = (m + (32u << lk) - 1) / (32u << lk), m2 = (m1 + 31) / 32;
  G1Xyzz* S0 = pool; pool += m1;
  G1Xyzz* W0 = pool; pool += m1;
  G1Xyzz* S1 = pool; pool += m2;
  G1Xyzz* W1 = pool; pool += m2;
  G1Xyzz* V = pool; pool += m2;
  if (ctx->opt_coop_tail) {
    G1Xyzz* res = pool; pool += 4;
    // level 1 stays one warp per chunk: all M/32 chains are resident at once, which four warps per chunk are not
    if (ctx->opt_coop_tail == 2) k_reduce_l1c<<<m1, 128, 0, ctx->stream>>>(cur, m, 1u << logT, W0, S0);
    else launch_l1(ctx, lk, dim3(m1, 1), cur, m, 1u << logT, W0, S0, 0);
    KZG_LAUNCHED(ctx);
    k_reduce_l2c<<<m2 * 2, 128, 0, ctx->stream>>>(S0, W0, m1, S1, W1, V, 0);
    KZG_LAUNCHED(ctx);
    k_reduce_l3c<<<3, 128, 0, ctx->stream>>>(S1, W1, V, m2, res, 0);
    KZG_LAUNCHED(ctx);
    k_reduce_horner<<<1, 128, 0, ctx->stream>>>(res, pre, (int)shallow, result_dev, 0, 5 + lk);
    KZG_LAUNCHED(ctx);
    return KZG_OK;
  }
  k_reduce_l1<<<m1, 32, 0, ctx->stream>>>(cur, m, 1u << logT, W0, S0, 0);
  k_reduce_l2<<<blocks_for((uint64_t)m2 * 64, 128), 128, 0, ctx->stream>>>(S0, W0, m1, S1, W1, V);
  KZG_LAUNCHED(ctx);
  k_reduce_l3<<<1, 96, 0, ctx->stream>>>(S1, W1, V, m2, pre, (int)shallow, result_dev);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

// B bucket sets of M <= 32768 buckets each (msm_bucket_multi): the fused cooperative tail once, blockIdx.y = set
int32_t reduce_buckets_multi(kzg_ctx* ctx, const G1Xyzz* parts, uint32_t M, uint32_t logT, int B, G1Xyzz* results_dev) {
  if (B == 1) return reduce_buckets(ctx, parts, M, logT, results_dev);
  const int lk = l1_chunk_log(ctx, M);
  if (M > (32768u << lk)) return fail(ctx, KZG_E_BAD_ARG, "fused bucket reduction: too many buckets per set");
  const uint32_t m1 = (M + (32u << lk) - 1) / (32u << lk), m2 = (m1 + 31) / 32;
  const uint32_t stride = 2 * m1 + 3 * m2 + 4;                 // S0 W0 | S1 W1 V | res, per set
  void* ws;
  KZG_TRY(scratch_get(ctx, SCR_XYZZ_TMP, ((size_t)stride * (size_t)B + 16) * sizeof(G1Xyzz), &ws));
  G1Xyzz* pool = reinterpret_cast<G1Xyzz*>(ws);
  G1Xyzz* S0 = pool; pool += m1;
  G1Xyzz* W0 = pool; pool += m1;
  G1Xyzz* S1 = pool; pool += m2;
  G1Xyzz* W1 = pool; pool += m2;
  G1Xyzz* V = pool; pool += m2;
  G1Xyzz* res = pool;
  const unsigned nb = (unsigned)B;
  launch_l1(ctx, lk, dim3(m1, nb), parts, M, 1u << logT, W0, S0, stride);
  KZG_LAUNCHED(ctx);
  k_reduce_l2c<<<dim3(m2 * 2, nb), 128, 0, ctx->stream>>>(S0, W0, m1, S1, W1, V, stride);
  KZG_LAUNCHED(ctx);
  k_reduce_l3c<<<dim3(3, nb), 128, 0, ctx->stream>>>(S1, W1, V, m2, res, stride);
  KZG_LAUNCHED(ctx);
  k_reduce_horner<<<dim3(1, nb), 128, 0, ctx->stream>>>(res, nullptr, 0, results_dev, stride, 5 + lk);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

// result[k] = 2^shift * res2[2k] + res2[2k + 1]: the hi and lo halves of a second-level bucket reduction (msm.cu)
__global__ void __launch_bounds__(128)
k_combine_l2(const G1Xyzz* __restrict__ res2, int shift, G1Xyzz* __restrict__ out) {
  __shared__ CoopBuf sh;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  G1Xyzz acc = res2[2 * blockIdx.x];
#pragma unroll 1
  for (int k = 0; k < shift; k++) dbl4(acc, &sh, w, lane);
  G1Xyzz lo = res2[2 * blockIdx.x + 1];
  add4(acc, lo, true, &sh, w, lane);
  if (threadIdx.x == 0) out[blockIdx.x] = acc;
}

int32_t combine_level2(kzg_ctx* ctx, const G1Xyzz* res2, int shift, int calls, G1Xyzz* results_dev) {
  k_combine_l2<<<(unsigned)calls, 128, 0, ctx->stream>>>(res2, shift, results_dev);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

}  // namespace kzg
