// K5' — fixed-base commits for LARGE batches: pairwise tree sums in affine coordinates with one shared
// inversion per round (Montgomery's trick across the whole grid).
//
// A blob commit through the fixed-base table (batch.cu) is a sum of n*W table points, all independent, all
// affine.  Accumulating them into an XYZZ register costs 8M + 2S = 10 Fq products per point.  Adding two AFFINE
// points costs one inversion + 2M + 1S; with the inversion shared by millions of independent additions
// (3 products each for the prefix-product / back-substitution of Montgomery's trick) an addition costs
// ~6.6 products.  So a batch is reduced in rounds: round 1 pairs up the W lookups of every coefficient, each
// later round pairs up neighbouring points of a row, every round halves the number of points:
//     k_*_den   : per pair the denominator d = x2 - x1 (2 y1 when the points are equal; 1 when an operand is
//                 infinity or the sum is) ; per thread the running product of its K denominators (prefixes to HBM)
//     batch_invert_inplace : products of 8, level by level, binary-GCD inversion of the few thousand values
//                 left at the top (field.cuh inv_gcd), and back
//     k_*_add   : d^-1 from the thread inverse and the prefixes; lambda, x3, y3; 96 B out
// and the last few hundred points per row are folded by an XYZZ kernel (k_rows_sum).  All additions are complete
// (infinity, P = Q, P = -Q) like blst_p1_add_or_double: setups with s = 0 / s = 1 and equal coefficients hit them.
// Replaces the same reference loop as batch.cu (src/polynomial.rs:207-212, :268) for rows >= tree_min_rows.
#include "affine_tree.cuh"

namespace kzg {

constexpr int TR_K = 16;           // pairs per thread in the generic rounds
constexpr int TR_THREADS = 128;
constexpr int INV_FAN = 8;         // fan-in of the batch-inversion product tree

// ---- round 1: the W table lookups of one coefficient, paired (2m, 2m+1) -----------------------------------
// signed digits of the scalar of (row, i): dig[j] = |d_j| | sign << 31
__device__ __forceinline__ void digits_of(const uint8_t* scalars, uint64_t row, uint64_t i, uint64_t stride, int c, int W,
                                          uint32_t* dig) {
  Fr k;
  const uint4* q = reinterpret_cast<const uint4*>(scalars + (row * stride + i) * 32);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  k.v[0] = a.x; k.v[1] = a.y; k.v[2] = a.z; k.v[3] = a.w;
  k.v[4] = b.x; k.v[5] = b.y; k.v[6] = b.z; k.v[7] = b.w;
  fr_canon_reduce(k);
  SignedDigits sd; sd.init(k, c);
  for (int j = 0; j < W; j++) {
    bool neg;
    const uint32_t d = sd.next(j, &neg);
    dig[j] = d | (neg ? 0x80000000u : 0u);
  }
}

__device__ __forceinline__ const G1Affine* tab_entry(const G1Affine* fix, uint64_t i, int W, int c, int j, uint32_t dg) {
  uint32_t d = dg & 0x7FFFFFFFu;
  if (j >= W || d == 0) return nullptr;
  return fix + ((((size_t)i * W + j) << (c - 1)) + (d - 1));
}

__global__ void __launch_bounds__(TR_THREADS)
k_r1_den(const uint8_t* __restrict__ scalars, uint64_t rows, uint64_t n, uint64_t stride, const G1Affine* __restrict__ fix,
         int c, int W, Fq* __restrict__ prefix, Fq* __restrict__ tot, int row_minor) {
  const uint64_t T = rows * n;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  // row_minor: neighbouring threads hold the SAME coefficient index of neighbouring rows, so a warp's lookups fall into
  // the two 2^(c-1)-entry sub-tables of one (coefficient, window pair) instead of 32 coefficients' worth of table
  const uint64_t row = row_minor ? t % rows : t / n, i = row_minor ? t / rows : t % n;
  uint32_t dig[32];
  digits_of(scalars, row, i, stride, c, W, dig);
  const int KP = (W + 1) / 2;
  Fq acc; set_one(acc);
#pragma unroll 1
  for (int m = 0; m < KP; m++) {
    const uint32_t d1 = dig[2 * m], d2 = (2 * m + 1 < W) ? dig[2 * m + 1] : 0u;
    const G1Affine* p1 = tab_entry(fix, i, W, c, 2 * m, d1);
    const G1Affine* p2 = tab_entry(fix, i, W, c, 2 * m + 1, d2);
    st_fq(prefix + (size_t)m * T + t, acc);
    if (p1 && p2) {               // the unpaired last lookup (odd W) and zero digits pass through: d = 1, no product
      Fq d;
      pair_den(d, p1, (d1 >> 31) != 0, p2, (d2 >> 31) != 0);
      mul(acc, acc, d);
    }
  }
  st_fq(tot + t, acc);
}

__global__ void __launch_bounds__(TR_THREADS)
k_r1_add(const uint8_t* __restrict__ scalars, uint64_t rows, uint64_t n, uint64_t stride, const G1Affine* __restrict__ fix,
         int c, int W, const Fq* __restrict__ prefix, const Fq* __restrict__ tot_inv, G1Affine* __restrict__ out, int row_minor) {
  const uint64_t T = rows * n;
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint64_t row = row_minor ? t % rows : t / n, i = row_minor ? t / rows : t % n;
  uint32_t dig[32];
  digits_of(scalars, row, i, stride, c, W, dig);
  const int KP = (W + 1) / 2;
  Fq inv; ld_fq_plain(inv, tot_inv + t);
#pragma unroll 1
  for (int m = KP - 1; m >= 0; m--) {
    const uint32_t d1 = dig[2 * m], d2 = (2 * m + 1 < W) ? dig[2 * m + 1] : 0u;
    const G1Affine* p1 = tab_entry(fix, i, W, c, 2 * m, d1);
    const G1Affine* p2 = tab_entry(fix, i, W, c, 2 * m + 1, d2);
    if (!p1 || !p2) {             // passes through (d = 1 in k_r1_den: the thread inverse is untouched)
      G1Affine a;
      ld_point(a, p1 ? p1 : p2, ((p1 ? d1 : d2) >> 31) != 0);
      G1Affine* dst = out + (size_t)(row * n + i) * KP + m;
      st_fq(&dst->x, a.x); st_fq(&dst->y, a.y);
      continue;
    }
    Fq d, pf, dinv;
    pair_den(d, p1, (d1 >> 31) != 0, p2, (d2 >> 31) != 0);
    ld_fq_plain(pf, prefix + (size_t)m * T + t);
    mul(dinv, inv, pf);
    mul(inv, inv, d);
    G1Affine a, b, o;
    ld_point(a, p1, (d1 >> 31) != 0);
    ld_point(b, p2, (d2 >> 31) != 0);
    pair_add(o, a, b, dinv);
    G1Affine* dst = out + (size_t)(row * n + i) * KP + m;      // rows stay contiguous for the generic rounds
    st_fq(&dst->x, o.x); st_fq(&dst->y, o.y);
  }
}

// ---- generic round: rows segments of length L -> length Lout = ceil(L/2) --------------------------------------
__global__ void __launch_bounds__(TR_THREADS)
k_rn_den(const G1Affine* __restrict__ in, uint64_t rows, uint64_t L, uint64_t Lout, uint64_t T,
         Fq* __restrict__ prefix, Fq* __restrict__ tot) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint64_t NP = rows * Lout;
  Fq acc; set_one(acc);
#pragma unroll 1
  for (int k = 0; k < TR_K; k++) {
    const uint64_t p = t + (uint64_t)k * T;
    if (p >= NP) break;
    const uint64_t row = p / Lout, q = p % Lout;
    const G1Affine* p1 = in + row * L + 2 * q;
    const G1Affine* p2 = (2 * q + 1 < L) ? p1 + 1 : nullptr;
    Fq d;
    pair_den(d, p1, false, p2, false);
    st_fq(prefix + (size_t)k * T + t, acc);
    mul(acc, acc, d);
  }
  st_fq(tot + t, acc);
}

__global__ void __launch_bounds__(TR_THREADS)
k_rn_add(const G1Affine* __restrict__ in, uint64_t rows, uint64_t L, uint64_t Lout, uint64_t T,
         const Fq* __restrict__ prefix, const Fq* __restrict__ tot_inv, G1Affine* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint64_t NP = rows * Lout;
  int kmax = 0;
  while (kmax < TR_K && t + (uint64_t)kmax * T < NP) kmax++;
  Fq inv; ld_fq_plain(inv, tot_inv + t);
#pragma unroll 1
  for (int k = kmax - 1; k >= 0; k--) {
    const uint64_t p = t + (uint64_t)k * T;
    const uint64_t row = p / Lout, q = p % Lout;
    const G1Affine* p1 = in + row * L + 2 * q;
    const G1Affine* p2 = (2 * q + 1 < L) ? p1 + 1 : nullptr;
    Fq d, pf, dinv;
    pair_den(d, p1, false, p2, false);
    ld_fq_plain(pf, prefix + (size_t)k * T + t);
    mul(dinv, inv, pf);
    mul(inv, inv, d);
    G1Affine a, b, o;
    ld_point(a, p1, false);
    ld_point(b, p2, false);
    pair_add(o, a, b, dinv);
    st_fq(&out[p].x, o.x); st_fq(&out[p].y, o.y);
  }
}

// ---- grid-wide batch inversion (in place) ----------------------------------------------------------------------
// level up: u < U owns v[u + k U], k < 8: pf[idx] = product of its earlier elements, up[u] = product of all
__global__ void __launch_bounds__(128)
k_inv_up(const Fq* __restrict__ v, uint64_t N, uint64_t U, Fq* __restrict__ pf, Fq* __restrict__ up) {
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  Fq acc; set_one(acc);
#pragma unroll 1
  for (int k = 0; k < INV_FAN; k++) {
    const uint64_t idx = u + (uint64_t)k * U;
    if (idx >= N) break;
    st_fq(pf + idx, acc);
    Fq x; ld_fq_plain(x, v + idx);
    mul(acc, acc, x);
  }
  st_fq(up + u, acc);
}
// level down: up[u] now holds the inverse of the product; v[idx] <- 1 / v[idx]
__global__ void __launch_bounds__(128)
k_inv_down(Fq* __restrict__ v, uint64_t N, uint64_t U, const Fq* __restrict__ pf, const Fq* __restrict__ up) {
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  int kmax = 0;
  while (kmax < INV_FAN && u + (uint64_t)kmax * U < N) kmax++;
  Fq inv; ld_fq_plain(inv, up + u);
#pragma unroll 1
  for (int k = kmax - 1; k >= 0; k--) {
    const uint64_t idx = u + (uint64_t)k * U;
    Fq x, p, r;
    ld_fq_plain(x, v + idx);
    ld_fq_plain(p, pf + idx);
    mul(r, inv, p);
    mul(inv, inv, x);
    st_fq(v + idx, r);
  }
}
__global__ void __launch_bounds__(64)
k_inv_top(Fq* __restrict__ v, uint64_t N) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  Fq x, r; ld_fq_plain(x, v + i);
  inv_gcd(r, x);          // uniform control flow: a warp of independent inversions does not diverge
  st_fq(v + i, r);
}

// top of the tree, fused: u < U owns v[u + k U], k < 8 — prefix products in registers, ONE binary-GCD inversion,
// back-substitution.  Replaces two levels + a per-value inversion kernel (five launches of pure latency).
template <int F>
__global__ void __launch_bounds__(64)
k_inv_top8(Fq* __restrict__ v, uint64_t N, uint64_t U) {
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u >= U) return;
  Fq x[F], pf[F];
  Fq acc; set_one(acc);
  int kmax = 0;
#pragma unroll
  for (int k = 0; k < F; k++) {
    const uint64_t idx = u + (uint64_t)k * U;
    if (idx < N) { ld_fq_plain(x[k], v + idx); pf[k] = acc; mul(acc, acc, x[k]); kmax = k + 1; }
  }
  Fq inv; inv_gcd(inv, acc);
#pragma unroll
  for (int k = F - 1; k >= 0; k--) {
    if (k < kmax) {
      Fq r; mul(r, inv, pf[k]);
      mul(inv, inv, x[k]);
      st_fq(v + u + (uint64_t)k * U, r);
    }
  }
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// scratch needed by batch_invert_inplace for N values, in Fq elements
size_t invert_scratch_elems(uint64_t N, uint64_t top) {
  size_t tot = 0;
  while (N > top) { uint64_t U = (N + INV_FAN - 1) / INV_FAN; tot += N + U; N = U; }
  return tot + 16;
}

// v[i] <- 1 / v[i], i < N (no zeros).  ws: invert_scratch_elems(N, top) Fq elements.
int32_t batch_invert_inplace(kzg_ctx* ctx, Fq* v, uint64_t N, Fq* ws, uint64_t top) {
  if (N == 0) return KZG_OK;
  struct Level { Fq* v; uint64_t N; uint64_t U; Fq* pf; Fq* up; };
  Level lv[24]; int nl = 0;
  Fq* cur = v; uint64_t n = N;
  while (n > top) {
    uint64_t U = (n + INV_FAN - 1) / INV_FAN;
    Level& l = lv[nl++];
    l.v = cur; l.N = n; l.U = U; l.pf = ws; ws += n; l.up = ws; ws += U;
    k_inv_up<<<blocks_for(U, 128), 128, 0, ctx->stream>>>(l.v, l.N, l.U, l.pf, l.up);
    KZG_LAUNCHED(ctx);
    cur = l.up; n = U;
  }
  if (n > 64) {
    const int fan = ctx->opt_inv_top_fan ? ctx->opt_inv_top_fan : (ctx->throughput_mode ? 8 : 4);
    if (fan == 4) {
      const uint64_t U = (n + 3) / 4;
      k_inv_top8<4><<<blocks_for(U, 64), 64, 0, ctx->stream>>>(cur, n, U);
    } else {
      const uint64_t U = (n + INV_FAN - 1) / INV_FAN;
      k_inv_top8<8><<<blocks_for(U, 64), 64, 0, ctx->stream>>>(cur, n, U);
    }
  } else {
    k_inv_top<<<blocks_for(n, 64), 64, 0, ctx->stream>>>(cur, n);
  }
  KZG_LAUNCHED(ctx);
  for (int i = nl - 1; i >= 0; i--) {
    k_inv_down<<<blocks_for(lv[i].U, 128), 128, 0, ctx->stream>>>(lv[i].v, lv[i].N, lv[i].U, lv[i].pf, lv[i].up);
    KZG_LAUNCHED(ctx);
  }
  return KZG_OK;
}

// ---- the last few hundred points of every row: XYZZ, one block per row ---------------------------------------------
__global__ void __launch_bounds__(128)
k_rows_sum(const G1Affine* __restrict__ in, uint64_t L, G1Xyzz* __restrict__ out) {
  __shared__ G1Xyzz sh[128];
  const uint64_t row = blockIdx.x;
  G1Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (uint64_t k = threadIdx.x; k < L; k += 128) {
    G1Affine p; ld_point(p, in + row * L + k, false);
    madd(acc, p, false);
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
#pragma unroll 1
  for (uint32_t st = 64; st >= 1; st >>= 1) {
    if (threadIdx.x < st) { G1Xyzz o = sh[threadIdx.x + st]; add(acc, o); sh[threadIdx.x] = acc; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[row] = acc;
}

int32_t msm_fixed_batch_tree(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t rows_total,
                             uint64_t n, uint64_t stride_scalars, G1Xyzz* out_dev) {
  if (rows_total == 0) return KZG_OK;
  if (!s->fix_tab || n > s->fix_n || n == 0) return fail(ctx, KZG_E_BAD_ARG, "fixed-base table not prepared for this size");
  const int c = s->fix_c, W = s->fix_w;
  if (W > 32) return fail(ctx, KZG_E_BAD_ARG, "tree path needs window_bits >= 8");
  const uint64_t KP = (uint64_t)(W + 1) / 2;
  const uint64_t chunk = (uint64_t)ctx->opt_tree_chunk_rows > 0 ? (uint64_t)ctx->opt_tree_chunk_rows : 2048;
  const uint64_t top = (uint64_t)ctx->opt_inv_top;   // values left for the fused top kernel (8 per thread)
  const uint64_t rows_max = rows_total < chunk ? rows_total : chunk;
  const uint64_t pts1 = rows_max * n * KP;             // points after round 1
  void *A, *B, *PF, *TW;
  KZG_TRY(scratch_get(ctx, SCR_TREE_A, pts1 * sizeof(G1Affine), &A));
  KZG_TRY(scratch_get(ctx, SCR_TREE_B, (pts1 / 2 + rows_max + 16) * sizeof(G1Affine), &B));
  KZG_TRY(scratch_get(ctx, SCR_TREE_PREFIX, pts1 * sizeof(Fq), &PF));
  const uint64_t T1 = rows_max * n;
  const uint64_t Tmax = T1 + rows_max + 16;          // >= threads of any round (KP <= 16 pairs per coefficient)
  KZG_TRY(scratch_get(ctx, SCR_TREE_TOT, (Tmax + invert_scratch_elems(Tmax, top)) * sizeof(Fq), &TW));
  Fq* prefix = reinterpret_cast<Fq*>(PF);
  Fq* tot = reinterpret_cast<Fq*>(TW);
  Fq* inv_ws = tot + Tmax;

  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  for (uint64_t r0 = 0; r0 < rows_total; r0 += chunk) {
    const uint64_t rows = rows_total - r0 < chunk ? rows_total - r0 : chunk;
    const uint8_t* sc = scalars_dev + r0 * stride_scalars * 32;
    // round 1
    const uint64_t T = rows * n;
    k_r1_den<<<blocks_for(T, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(sc, rows, n, stride_scalars, s->fix_tab, c, W, prefix, tot, ctx->opt_r1_row_minor);
    KZG_LAUNCHED(ctx);
    KZG_TRY(batch_invert_inplace(ctx, tot, T, inv_ws, top));
    G1Affine* cur = reinterpret_cast<G1Affine*>(A);
    G1Affine* nxt = reinterpret_cast<G1Affine*>(B);
    KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[0], ctx->stream));      // the dominant kernel of a blob batch (kzg_ctx_last_kernel_ms(2))
    k_r1_add<<<blocks_for(T, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(sc, rows, n, stride_scalars, s->fix_tab, c, W, prefix, tot, cur, ctx->opt_r1_row_minor);
    KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[1], ctx->stream));
    KZG_LAUNCHED(ctx);
    uint64_t L = n * KP;
    // generic rounds while there is enough independent work to hide the inversion latency
    while (L > 1) {
      const uint64_t Lout = (L + 1) / 2;
      const uint64_t NP = rows * Lout;
      if (NP < (uint64_t)ctx->opt_tree_min_pairs) break;
      const uint64_t Tn = (NP + TR_K - 1) / TR_K;
      k_rn_den<<<blocks_for(Tn, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(cur, rows, L, Lout, Tn, prefix, tot);
      KZG_LAUNCHED(ctx);
      KZG_TRY(batch_invert_inplace(ctx, tot, Tn, inv_ws, top));
      k_rn_add<<<blocks_for(Tn, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(cur, rows, L, Lout, Tn, prefix, tot, nxt);
      KZG_LAUNCHED(ctx);
      G1Affine* t = cur; cur = nxt; nxt = t;
      L = Lout;
    }
    k_rows_sum<<<(unsigned)rows, 128, 0, ctx->stream>>>(cur, L, out_dev + r0);
    KZG_LAUNCHED(ctx);
  }
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  return KZG_OK;
}

// ---- N fixed-base multiplications of ONE base (powers-of-tau generation, setup.cu) ------------------------------------
// out[i] = k_i * B with tab[(j << (c-1)) + d-1] = d * 2^(c j) * B: every output is the sum of W table points, which is
// exactly a "row" of one coefficient for the kernels above (n = 1, stride = 1, base index 0).  Round 1 pairs the W
// lookups, log2(W/2) generic rounds finish the sum, and the last round writes the AFFINE result straight into `out`:
// 15 additions x 6 products = 90 Fq products per point and no per-point inversion, against 16 mixed additions x 10 +
// the batched to-affine (~165) on the XYZZ path (k_setup_points).
int32_t fixed_base_points_tree(kzg_ctx* ctx, const G1Affine* tab, int c, int W, const uint8_t* scalars_dev, uint64_t N,
                               G1Affine* out_dev) {
  if (N == 0) return KZG_OK;
  if (W > 32 || W < 3) return fail(ctx, KZG_E_BAD_ARG, "tree path needs 3..32 windows");
  const uint64_t KP = (uint64_t)(W + 1) / 2;
  const uint64_t chunk = 1ull << 22;   // rows per pass: each of the 4 rounds pays one ~0.2 ms inversion chain (6.5 GB of scratch)
  const uint64_t top = (uint64_t)ctx->opt_inv_top;
  const uint64_t rows_max = N < chunk ? N : chunk;
  const uint64_t pts1 = rows_max * KP;
  void *A, *B, *PF, *TW;
  KZG_TRY(scratch_get(ctx, SCR_TREE_A, pts1 * sizeof(G1Affine), &A));
  KZG_TRY(scratch_get(ctx, SCR_TREE_B, (pts1 / 2 + rows_max + 16) * sizeof(G1Affine), &B));
  KZG_TRY(scratch_get(ctx, SCR_TREE_PREFIX, pts1 * sizeof(Fq), &PF));
  const uint64_t Tmax = rows_max + 16;
  KZG_TRY(scratch_get(ctx, SCR_TREE_TOT, (Tmax + invert_scratch_elems(Tmax, top)) * sizeof(Fq), &TW));
  Fq* prefix = reinterpret_cast<Fq*>(PF);
  Fq* tot = reinterpret_cast<Fq*>(TW);
  Fq* inv_ws = tot + Tmax;
  for (uint64_t r0 = 0; r0 < N; r0 += chunk) {
    const uint64_t rows = N - r0 < chunk ? N - r0 : chunk;
    const uint8_t* sc = scalars_dev + r0 * 32;
    k_r1_den<<<blocks_for(rows, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(sc, rows, 1, 1, tab, c, W, prefix, tot, 0);
    KZG_LAUNCHED(ctx);
    KZG_TRY(batch_invert_inplace(ctx, tot, rows, inv_ws, top));
    G1Affine* cur = reinterpret_cast<G1Affine*>(A);
    G1Affine* nxt = reinterpret_cast<G1Affine*>(B);
    k_r1_add<<<blocks_for(rows, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(sc, rows, 1, 1, tab, c, W, prefix, tot, cur, 0);
    KZG_LAUNCHED(ctx);
    uint64_t L = KP;
    while (L > 1) {
      const uint64_t Lout = (L + 1) / 2;
      const uint64_t NP = rows * Lout;
      const uint64_t Tn = (NP + TR_K - 1) / TR_K;
      G1Affine* dst = Lout == 1 ? out_dev + r0 : nxt;
      k_rn_den<<<blocks_for(Tn, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(cur, rows, L, Lout, Tn, prefix, tot);
      KZG_LAUNCHED(ctx);
      KZG_TRY(batch_invert_inplace(ctx, tot, Tn, inv_ws, top));
      k_rn_add<<<blocks_for(Tn, TR_THREADS), TR_THREADS, 0, ctx->stream>>>(cur, rows, L, Lout, Tn, prefix, tot, dst);
      KZG_LAUNCHED(ctx);
      nxt = cur; cur = dst;
      L = Lout;
    }
  }
  return KZG_OK;
}

}  // namespace kzg
