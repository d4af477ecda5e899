// K3' — the G2 half of SetupArtifact: item i carries [s^i]G2 next to [s^i]G1
// (src/trusted_setup.rs:31-35, 64-72: one 256-bit blst_p2_mult on the G2 generator per item; it dominates the
// reference's setup time).  Same structure as the G1 side: the running power s^i comes from chunked prefix products,
// the scalar multiplication becomes a fixed-base table walk over SIGNED 13-bit windows of the scalar folded to
// min(k, r - k) — 20 lookups into T2[j][d-1] = d * 2^(13 j) * G2 (20 x 4096 affine points over Fq2, 15 MiB,
// L2-resident) — and each point leaves in its canonical form: the 96-byte compressed encoding of Serialize for
// G2Point (src/curves.rs:260-271).
//
// Two ways to add the 20 table points of an item:
//   * small ranges: one thread per item, XYZZ += affine over Fq2 (28 Fq products per addition) and one Fq2 inversion
//     per item.  255 registers and spills, 2 warps per scheduler: 35 M points/s on large ranges.
//   * ranges >= setup_tree_min: affine pairwise trees with grid-wide shared inversions, like the G1 side
//     (batch_tree.cu).  The denominator d = x2 - x1 lives in Fq2, but 1/d = conj(d) / N(d) with the norm
//     N(d) = d0^2 + d1^2 in Fq, so the shared inversion is the existing Fq one (batch_invert_inplace) over the norms:
//     2 S for the norm + 3 M for Montgomery's trick + 2 M for conj(d)/N, then lambda (3), lambda^2 (2), y3 (3) =
//     15 Fq products per addition, 19 additions per item, no per-item inversion: 55 M points/s.
// All additions are complete (infinity, P = Q, P = -Q): s = 0 / s = 1 setups and cancelling digits hit them.
#include "affine_tree.cuh"
#include "g2.cuh"

namespace kzg {

constexpr int G2T_C = 13;                 // signed window bits
constexpr int G2T_W = 20;                 // windows_for(13) = ceil(255 / 13)
constexpr int G2T_HALF = 1 << (G2T_C - 1);   // multiples per window
constexpr int G2T_KP = G2T_W / 2;         // pairs per item in round 1
static_assert(G2T_W * G2T_C >= 255 && (G2T_W % 2) == 0, "window geometry");

__device__ __forceinline__ void inv_dev(Fq2& o, const Fq2& a) {
  Fq n, t;
  sqr(n, a.c0); sqr(t, a.c1); add(n, n, t);
  inv_gcd(t, n);          // uniform control flow (field.cuh), ~12x fewer instructions than Fermat
  mul(o.c0, a.c0, t);
  mul(t, a.c1, t);
  neg(o.c1, t);
}

// T2[j*HALF + d-1] = d * 2^(13 j) * G2 : thread = one table entry (built once per context)
__global__ void __launch_bounds__(64)
k_g2_table(G2Affine* __restrict__ tab) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G2T_W * G2T_HALF) return;
  uint32_t j = t / G2T_HALF, d = t % G2T_HALF + 1;
  uint32_t k[10];
#pragma unroll
  for (int i = 0; i < 10; i++) k[i] = 0;
  const uint32_t bit = j * G2T_C;
  const uint64_t sh = (uint64_t)d << (bit & 31);          // d < 2^13: at most two limbs
  k[bit >> 5] = (uint32_t)sh; k[(bit >> 5) + 1] = (uint32_t)(sh >> 32);
  G2Affine g; g2_generator(g);
  G2Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (int i = (int)bit + G2T_C - 1; i >= 0; i--) {
    dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) madd(acc, g.x, g.y);
  }
  G2Affine a;
  if (is_inf(acc)) set_inf(a);
  else { Fq2 w; inv_dev(w, acc.zzz); to_affine_with_inv(a, acc, w); }
  tab[t] = a;
}

// signed digits of a canonical scalar, folded to min(k, r - k): dig[j] = |d_j| | sign << 31 (recode.cuh)
__device__ __forceinline__ void g2_digits(const Fr& kin, uint32_t* dig) {
  SignedDigits sd; sd.init(kin, G2T_C);
#pragma unroll 1
  for (int j = 0; j < G2T_W; j++) {
    bool neg;
    const uint32_t d = sd.next(j, &neg);
    dig[j] = d | (neg ? 0x80000000u : 0u);
  }
}

__device__ __forceinline__ void neg_y(G2Affine& p) { neg(p.y, p.y); }

__global__ void __launch_bounds__(64)
k_setup_points_g2(const Fr* __restrict__ scalars, uint64_t n, const G2Affine* __restrict__ tab,
                  uint8_t* __restrict__ out_c96) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  SignedDigits sd; sd.init(scalars[i], G2T_C);
  G2Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (int j = 0; j < G2T_W; j++) {
    bool neg;
    const uint32_t d = sd.next(j, &neg);
    if (d != 0) {
      G2Affine p = tab[(size_t)j * G2T_HALF + (d - 1)];
      if (neg) neg_y(p);
      madd(acc, p.x, p.y);
    }
  }
  G2Affine a;
  if (is_inf(acc)) set_inf(a);
  else { Fq2 w; inv_dev(w, acc.zzz); to_affine_with_inv(a, acc, w); }
  compress(out_c96 + i * 96, a);
}

__global__ void k_powers_g2(Fr s_mont, uint64_t first, uint64_t n, uint32_t K, Fr* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i0 = t * K;
  if (i0 >= n) return;
  Fr cur; pow_u64(cur, s_mont, first + i0);
  Fr c; from_mont(c, cur);
  for (uint32_t k = 0; k < K && i0 + k < n; k++) {
    out[i0 + k] = c;
    Fr nx; mul(nx, s_mont, c);
    c = nx;
  }
}

// ---- affine pairwise trees over Fq2 -------------------------------------------------------------------------------
__device__ __forceinline__ void ld_fq2(Fq2& o, const Fq2* p) { ld_fq_plain(o.c0, &p->c0); ld_fq_plain(o.c1, &p->c1); }
__device__ __forceinline__ void st_fq2(Fq2* p, const Fq2& v) { st_fq(&p->c0, v.c0); st_fq(&p->c1, v.c1); }

__device__ __forceinline__ const G2Affine* g2_entry(const G2Affine* tab, int j, uint32_t dg) {
  const uint32_t d = dg & 0x7FFFFFFFu;
  return d ? tab + ((size_t)j * G2T_HALF + (d - 1)) : nullptr;
}

// Denominator of P1 + P2 and its norm.  d = x2 - x1; 2 y1 when P1 = P2; 1 when an operand or the sum is infinity.
// p == nullptr stands for infinity; y is only touched on the rare paths.
__device__ __forceinline__ void g2_pair_den(Fq2& d, Fq& nrm, const G2Affine* p1, bool n1, const G2Affine* p2, bool n2) {
  set_one(d);
  bool plain = false;
  if (p1 && p2) {
    Fq2 x1, x2;
    ld_fq2(x1, &p1->x); ld_fq2(x2, &p2->x);
    bool inf = false;
    if (is_zero(x1) || is_zero(x2)) {
      Fq2 y1, y2;
      ld_fq2(y1, &p1->y); ld_fq2(y2, &p2->y);
      inf = (is_zero(x1) && is_zero(y1)) || (is_zero(x2) && is_zero(y2));
    }
    if (!inf) {
      Fq2 t; sub(t, x2, x1);
      if (!is_zero(t)) { d = t; plain = true; }
      else {
        Fq2 y1, y2;
        ld_fq2(y1, &p1->y); ld_fq2(y2, &p2->y);
        if (n1) neg(y1, y1);
        if (n2) neg(y2, y2);
        if (eq(y1, y2) && !is_zero(y1)) { dbl(d, y1); plain = true; }
      }
    }
  }
  if (plain) { Fq t; sqr(nrm, d.c0); sqr(t, d.c1); add(nrm, nrm, t); }     // N(d) != 0 for d != 0: -1 is a non-residue
  else set_one(nrm);
}

__device__ __forceinline__ void g2_ld_point(G2Affine& o, const G2Affine* p, bool negate) {
  if (!p) { set_inf(o); return; }
  ld_fq2(o.x, &p->x); ld_fq2(o.y, &p->y);
  if (negate && !is_inf(o)) neg(o.y, o.y);
}

// P1 + P2 given dinv = 1 / g2_pair_den(P1, P2); the case analysis mirrors g2_pair_den exactly.
__device__ __forceinline__ void g2_pair_add(G2Affine& o, const G2Affine& a, const G2Affine& b, const Fq2& dinv) {
  if (is_inf(a)) { o = b; return; }
  if (is_inf(b)) { o = a; return; }
  Fq2 lam, t;
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
  Fq2 x3;
  sqr(x3, lam); sub(x3, x3, a.x); sub(x3, x3, b.x);
  sub(t, a.x, x3); mul(t, lam, t); sub(o.y, t, a.y);
  o.x = x3;
}

// 1/d from the inverse of its norm: conj(d) * ninv
__device__ __forceinline__ void g2_dinv(Fq2& o, const Fq2& d, const Fq& ninv) {
  Fq t;
  mul(o.c0, d.c0, ninv);
  mul(t, d.c1, ninv);
  neg(o.c1, t);
}

constexpr int G2_TR_K = 8;            // pairs per thread in the generic rounds

// round 1: the 20 table lookups of one item, paired (2m, 2m+1)
__global__ void __launch_bounds__(128)
k_g2_r1_den(const Fr* __restrict__ scalars, uint64_t T, const G2Affine* __restrict__ tab,
            Fq* __restrict__ prefix, Fq* __restrict__ tot) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint32_t dig[G2T_W];
  g2_digits(scalars[t], dig);
  Fq acc; set_one(acc);
#pragma unroll 1
  for (int m = 0; m < G2T_KP; m++) {
    const uint32_t d1 = dig[2 * m], d2 = dig[2 * m + 1];
    Fq2 d; Fq nrm;
    g2_pair_den(d, nrm, g2_entry(tab, 2 * m, d1), (d1 >> 31) != 0, g2_entry(tab, 2 * m + 1, d2), (d2 >> 31) != 0);
    st_fq(prefix + (size_t)m * T + t, acc);
    mul(acc, acc, nrm);
  }
  st_fq(tot + t, acc);
}

__global__ void __launch_bounds__(128)
k_g2_r1_add(const Fr* __restrict__ scalars, uint64_t T, const G2Affine* __restrict__ tab,
            const Fq* __restrict__ prefix, const Fq* __restrict__ tot_inv, G2Affine* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  uint32_t dig[G2T_W];
  g2_digits(scalars[t], dig);
  Fq inv; ld_fq_plain(inv, tot_inv + t);
#pragma unroll 1
  for (int m = G2T_KP - 1; m >= 0; m--) {
    const uint32_t d1 = dig[2 * m], d2 = dig[2 * m + 1];
    const G2Affine* p1 = g2_entry(tab, 2 * m, d1);
    const G2Affine* p2 = g2_entry(tab, 2 * m + 1, d2);
    Fq2 d, dinv; Fq nrm, pf, ninv;
    g2_pair_den(d, nrm, p1, (d1 >> 31) != 0, p2, (d2 >> 31) != 0);
    ld_fq_plain(pf, prefix + (size_t)m * T + t);
    mul(ninv, inv, pf);
    mul(inv, inv, nrm);
    g2_dinv(dinv, d, ninv);
    G2Affine a, b, o;
    g2_ld_point(a, p1, (d1 >> 31) != 0);
    g2_ld_point(b, p2, (d2 >> 31) != 0);
    g2_pair_add(o, a, b, dinv);
    G2Affine* dst = out + (size_t)t * G2T_KP + m;
    st_fq2(&dst->x, o.x); st_fq2(&dst->y, o.y);
  }
}

// generic round: rows segments of length L -> length Lout = ceil(L / 2)
__global__ void __launch_bounds__(128)
k_g2_rn_den(const G2Affine* __restrict__ in, uint64_t rows, uint64_t L, uint64_t Lout, uint64_t T,
            Fq* __restrict__ prefix, Fq* __restrict__ tot) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint64_t NP = rows * Lout;
  Fq acc; set_one(acc);
#pragma unroll 1
  for (int k = 0; k < G2_TR_K; k++) {
    const uint64_t p = t + (uint64_t)k * T;
    if (p >= NP) break;
    const uint64_t row = p / Lout, q = p % Lout;
    const G2Affine* p1 = in + row * L + 2 * q;
    const G2Affine* p2 = (2 * q + 1 < L) ? p1 + 1 : nullptr;
    Fq2 d; Fq nrm;
    g2_pair_den(d, nrm, p1, false, p2, false);
    st_fq(prefix + (size_t)k * T + t, acc);
    mul(acc, acc, nrm);
  }
  st_fq(tot + t, acc);
}

__global__ void __launch_bounds__(128)
k_g2_rn_add(const G2Affine* __restrict__ in, uint64_t rows, uint64_t L, uint64_t Lout, uint64_t T,
            const Fq* __restrict__ prefix, const Fq* __restrict__ tot_inv, G2Affine* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T) return;
  const uint64_t NP = rows * Lout;
  int kmax = 0;
  while (kmax < G2_TR_K && t + (uint64_t)kmax * T < NP) kmax++;
  Fq inv; ld_fq_plain(inv, tot_inv + t);
#pragma unroll 1
  for (int k = kmax - 1; k >= 0; k--) {
    const uint64_t p = t + (uint64_t)k * T;
    const uint64_t row = p / Lout, q = p % Lout;
    const G2Affine* p1 = in + row * L + 2 * q;
    const G2Affine* p2 = (2 * q + 1 < L) ? p1 + 1 : nullptr;
    Fq2 d, dinv; Fq nrm, pf, ninv;
    g2_pair_den(d, nrm, p1, false, p2, false);
    ld_fq_plain(pf, prefix + (size_t)k * T + t);
    mul(ninv, inv, pf);
    mul(inv, inv, nrm);
    g2_dinv(dinv, d, ninv);
    G2Affine a, b, o;
    g2_ld_point(a, p1, false);
    g2_ld_point(b, p2, false);
    g2_pair_add(o, a, b, dinv);
    st_fq2(&out[p].x, o.x); st_fq2(&out[p].y, o.y);
  }
}

__global__ void __launch_bounds__(128)
k_g2_compress(const G2Affine* __restrict__ in, uint64_t n, uint8_t* __restrict__ out_c96) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G2Affine a; g2_ld_point(a, in + i, false);
  compress(out_c96 + i * 96, a);
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// N items through the trees, in passes of `chunk` items: round 1 (20 -> 10 points per item), then 10 -> 5 -> 3 -> 2 -> 1
static int32_t g2_points_tree(kzg_ctx* ctx, const G2Affine* tab, const Fr* scalars_dev, uint64_t N, uint8_t* out_c96_dev) {
  const uint64_t chunk = 1ull << 20;                     // 2 GB + 1 GB of intermediate points per pass
  const uint64_t top = (uint64_t)ctx->opt_inv_top;
  const uint64_t rows_max = N < chunk ? N : chunk;
  const uint64_t pts1 = rows_max * G2T_KP;
  void *A, *B, *PF, *TW;
  KZG_TRY(scratch_get(ctx, SCR_TREE_A, pts1 * sizeof(G2Affine), &A));
  KZG_TRY(scratch_get(ctx, SCR_TREE_B, (pts1 / 2 + rows_max + 16) * sizeof(G2Affine), &B));
  KZG_TRY(scratch_get(ctx, SCR_TREE_PREFIX, pts1 * sizeof(Fq), &PF));
  const uint64_t Tmax = rows_max + 16;
  KZG_TRY(scratch_get(ctx, SCR_TREE_TOT, (Tmax + invert_scratch_elems(Tmax, top)) * sizeof(Fq), &TW));
  Fq* prefix = reinterpret_cast<Fq*>(PF);
  Fq* tot = reinterpret_cast<Fq*>(TW);
  Fq* inv_ws = tot + Tmax;
  for (uint64_t r0 = 0; r0 < N; r0 += chunk) {
    const uint64_t rows = N - r0 < chunk ? N - r0 : chunk;
    const Fr* sc = scalars_dev + r0;
    k_g2_r1_den<<<blocks_for(rows, 128), 128, 0, ctx->stream>>>(sc, rows, tab, prefix, tot);
    KZG_LAUNCHED(ctx);
    KZG_TRY(batch_invert_inplace(ctx, tot, rows, inv_ws, top));
    G2Affine* cur = reinterpret_cast<G2Affine*>(A);
    G2Affine* nxt = reinterpret_cast<G2Affine*>(B);
    k_g2_r1_add<<<blocks_for(rows, 128), 128, 0, ctx->stream>>>(sc, rows, tab, prefix, tot, cur);
    KZG_LAUNCHED(ctx);
    uint64_t L = G2T_KP;
    while (L > 1) {
      const uint64_t Lout = (L + 1) / 2;
      const uint64_t NP = rows * Lout;
      const uint64_t Tn = (NP + G2_TR_K - 1) / G2_TR_K;
      k_g2_rn_den<<<blocks_for(Tn, 128), 128, 0, ctx->stream>>>(cur, rows, L, Lout, Tn, prefix, tot);
      KZG_LAUNCHED(ctx);
      KZG_TRY(batch_invert_inplace(ctx, tot, Tn, inv_ws, top));
      k_g2_rn_add<<<blocks_for(Tn, 128), 128, 0, ctx->stream>>>(cur, rows, L, Lout, Tn, prefix, tot, nxt);
      KZG_LAUNCHED(ctx);
      G2Affine* t = cur; cur = nxt; nxt = t;
      L = Lout;
    }
    k_g2_compress<<<blocks_for(rows, 128), 128, 0, ctx->stream>>>(cur, rows, out_c96_dev + r0 * 96);
    KZG_LAUNCHED(ctx);
  }
  return KZG_OK;
}

int32_t setup_generate_g2_dev(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n,
                              uint8_t* out_c96_dev) {
  if (n == 0) return KZG_OK;
  if (!ctx->g2table) {
    G2Affine* tab = nullptr;
    KZG_CUDA_OK(ctx, cudaMalloc(&tab, sizeof(G2Affine) * G2T_W * G2T_HALF));
    k_g2_table<<<blocks_for(G2T_W * G2T_HALF, 64), 64, 0, ctx->stream>>>(tab);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(tab); ctx->err = std::string("k_g2_table: ") + cudaGetErrorString(e); return KZG_E_CUDA; }
    ctx->g2table = tab;
  }
  uint8_t le[32];
  for (int i = 0; i < 32; i++) le[i] = secret_be[31 - i];
  Fr raw, s_mont; fr_load_le(raw, le); to_mont(s_mont, raw);
  void* pw;
  KZG_TRY(scratch_get(ctx, SCR_FR_B, n * sizeof(Fr), &pw));
  const uint32_t K = 32;
  k_powers_g2<<<blocks_for((n + K - 1) / K, 128), 128, 0, ctx->stream>>>(s_mont, first, n, K, reinterpret_cast<Fr*>(pw));
  KZG_LAUNCHED(ctx);
  const G2Affine* tab = reinterpret_cast<const G2Affine*>(ctx->g2table);
  // the G2 trees pay off earlier than the G1 ones (the XYZZ walk runs at 2 warps per scheduler): a quarter of the G1
  // threshold, 2^16 items by default (measured: 2^12 0.91 ms walk / 1.36 trees, 2^18 10.0 / 6.7, 2^20 30.0 / 18.9)
  if (ctx->opt_setup_tree_min > 0 && n >= (uint64_t)std::max(1, ctx->opt_setup_tree_min / 4))
    return g2_points_tree(ctx, tab, reinterpret_cast<const Fr*>(pw), n, out_c96_dev);
  k_setup_points_g2<<<blocks_for(n, 64), 64, 0, ctx->stream>>>(reinterpret_cast<Fr*>(pw), n, tab, out_c96_dev);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

}  // namespace kzg
