// K3 — powers-of-tau generation and per-setup table precomputation.
//
// Replaces SetupArtifactsGenerator::next (src/trusted_setup.rs:40-78): the running product
// current_s_powered * secret (:50) becomes a chunked prefix-power kernel, and the 256-bit
// blst_p1_mult on the generator (:54-62) becomes a fixed-base table walk: 16 signed 16-bit
// windows, table T[j][d-1] = d * 2^(16 j) * G (16 x 32768 affine points, 48 MiB, L2-resident),
// so one setup point costs 16 mixed additions.  XYZZ results are converted to affine with one
// shared inversion per thread (Montgomery's trick) and kept device-resident.
// Also here: G1 (de)serialisation kernels (Serialize / Deserialize for G1Point,
// src/curves.rs:99-183) and the per-setup tables used by the MSM kernels.
#include <cmath>
#include "engine.cuh"

namespace kzg {

constexpr int GT_C = 16;            // generator-table window bits
constexpr int GT_W = 16;            // windows
constexpr int GT_HALF = 1 << 15;    // multiples per window

__device__ __forceinline__ void ld_affine(G1Affine& o, const G1Affine* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 r[6];
#pragma unroll
  for (int i = 0; i < 6; i++) r[i] = __ldg(q + i);
#pragma unroll
  for (int i = 0; i < 3; i++) {
    o.x.v[4 * i] = r[i].x; o.x.v[4 * i + 1] = r[i].y; o.x.v[4 * i + 2] = r[i].z; o.x.v[4 * i + 3] = r[i].w;
    o.y.v[4 * i] = r[3 + i].x; o.y.v[4 * i + 1] = r[3 + i].y; o.y.v[4 * i + 2] = r[3 + i].z; o.y.v[4 * i + 3] = r[3 + i].w;
  }
}

// ---- powers of s -----------------------------------------------------------------------
// out[i] = s^(first + i) as a canonical integer, i < n; thread t owns K consecutive items.
__global__ void k_powers(Fr s_mont, uint64_t first, uint64_t n, uint32_t K, Fr* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i0 = t * K;
  if (i0 >= n) return;
  Fr cur; pow_u64(cur, s_mont, first + i0);
  Fr c; from_mont(c, cur);                     // canonical
  for (uint32_t k = 0; k < K && i0 + k < n; k++) {
    out[i0 + k] = c;
    Fr nx; mul(nx, s_mont, c);                 // (sR) * c * R^-1 = s*c, stays canonical
    c = nx;
  }
}

// ---- fixed-base generator multiplication -------------------------------------------------
__global__ void __launch_bounds__(128)
k_setup_points(const Fr* __restrict__ scalars, uint64_t n, const G1Affine* __restrict__ gtable,
               G1Xyzz* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = scalars[i];
  G1Xyzz acc; set_inf(acc);
  SignedDigits sd; sd.init(k, GT_C);
#pragma unroll 1
  for (int j = 0; j < GT_W; j++) {
    bool neg;
    const uint32_t d = sd.next(j, &neg);
    if (d != 0) {
      G1Affine p; ld_affine(p, gtable + (size_t)j * GT_HALF + (d - 1));
      madd(acc, p, neg);
    }
  }
  out[i] = acc;
}

// ---- multiples table fill: out[b*half + d-1] = d * B_b, thread = (b, chunk of K) ---------------
// B_b is read from tab[(b % W) * n_stride + (b / W)]  (b = i*W + j over a folded-window table)
__global__ void __launch_bounds__(128)
k_fill_multiples(const G1Affine* __restrict__ tab, uint64_t n_stride, uint32_t W, uint64_t b_first,
                 uint64_t b_count, uint32_t half, uint32_t K, G1Xyzz* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t chunks = half / K;
  uint64_t bl = t / chunks;
  uint32_t ch = (uint32_t)(t % chunks);
  if (bl >= b_count) return;
  uint64_t b = b_first + bl;
  G1Affine B; ld_affine(B, tab + (size_t)(b % W) * n_stride + (b / W));
  G1Xyzz* o = out + (size_t)bl * half + (size_t)ch * K;
  if (is_inf(B)) {
    G1Xyzz inf; set_inf(inf);
    for (uint32_t k = 0; k < K; k++) o[k] = inf;
    return;
  }
  uint32_t d0 = ch * K + 1;
  G1Xyzz acc; scalar_mul(acc, B, &d0, 32 - __clz(d0));
  o[0] = acc;
  for (uint32_t k = 1; k < K; k++) { madd(acc, B.x, B.y); o[k] = acc; }
}

// ---- out[i] = 2^c * in[i] ---------------------------------------------------------------------
__global__ void __launch_bounds__(128)
k_dbl_c(const G1Affine* __restrict__ in, uint64_t n, int c, G1Xyzz* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p; ld_affine(p, in + i);
  G1Xyzz acc;
  if (is_inf(p)) { set_inf(acc); out[i] = acc; return; }
  dbl_affine(acc, p.x, p.y);
  for (int k = 1; k < c; k++) dbl(acc);
  out[i] = acc;
}

// ---- batch XYZZ -> affine (one inversion per thread, strided ownership for coalescing) ---------
__global__ void __launch_bounds__(128)
k_batch_affine(const G1Xyzz* __restrict__ in, uint64_t n, G1Affine* __restrict__ out,
               Fq* __restrict__ prefix, uint64_t T) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= T || t >= n) return;
  Fq acc; set_one(acc);
  uint64_t last = t;
  for (uint64_t i = t; i < n; i += T) {
    prefix[i] = acc;
    if (!is_zero(in[i].zz)) { Fq z = in[i].zzz; mul(acc, acc, z); }
    last = i;
  }
  Fq ai; inv_gcd(ai, acc);
  for (uint64_t i = last;; i -= T) {
    G1Xyzz p = in[i];
    if (is_inf(p)) {
      G1Affine a; set_inf(a); out[i] = a;
    } else {
      Fq w; mul(w, ai, prefix[i]);
      mul(ai, ai, p.zzz);
      G1Affine a; to_affine_with_inv(a, p, w);
      out[i] = a;
    }
    if (i < T + t || i == t) break;
  }
}

// ---- (de)serialisation ---------------------------------------------------------------------------
__global__ void k_compress(const G1Affine* __restrict__ in, uint64_t n, uint8_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Affine p; ld_affine(p, in + i);
  uint32_t w[12];
  if (is_inf(p)) {
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = 0;
    w[0] = 0xC0;
  } else {
    Fq x, y; from_mont(x, p.x); from_mont(y, p.y);
#pragma unroll
    for (int k = 0; k < 12; k++) w[k] = __byte_perm(x.v[11 - k], 0, 0x0123);
    w[0] |= 0x80;
    if (fq_canon_gt_half(y)) w[0] |= 0x20;
  }
  uint4* o = reinterpret_cast<uint4*>(out + i * 48);
  o[0] = make_uint4(w[0], w[1], w[2], w[3]);
  o[1] = make_uint4(w[4], w[5], w[6], w[7]);
  o[2] = make_uint4(w[8], w[9], w[10], w[11]);
}

__global__ void k_compress_xyzz(const G1Xyzz* __restrict__ in, uint64_t n, uint8_t* __restrict__ out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  G1Xyzz p = in[i];
  G1Affine a; to_affine(a, p);
  uint8_t buf[48];
  compress(buf, a);
  for (int k = 0; k < 48; k++) out[i * 48 + k] = buf[k];
}

__global__ void k_uncompress(const uint8_t* __restrict__ in, uint64_t n, G1Affine* __restrict__ out,
                             unsigned long long* __restrict__ bad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint8_t buf[48];
  for (int k = 0; k < 48; k++) buf[k] = in[i * 48 + k];
  G1Affine a;
  if (uncompress(a, buf)) { atomicMin(bad, (unsigned long long)i); set_inf(a); }
  out[i] = a;
}

// dense [n] x 96 B  ->  [n] x 128 B pitch (one point per DRAM line); 8 threads per point, 16 bytes each
__global__ void k_pitch_table(const uint4* __restrict__ in, uint64_t n, uint4* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i = t >> 3; uint32_t q = (uint32_t)t & 7u;
  if (i >= n) return;
  out[i * 8 + q] = q < 6 ? in[i * 6 + q] : make_uint4(0, 0, 0, 0);
}

__global__ void k_set_generator(G1Affine* out) { G1Affine g; g1_generator(g); out[0] = g; }

// -----------------------------------------------------------------------------------------------
static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int32_t batch_to_affine(kzg_ctx* ctx, const G1Xyzz* in_dev, uint64_t n, G1Affine* out_dev) {
  if (n == 0) return KZG_OK;
  void* pf;
  KZG_TRY(scratch_get(ctx, SCR_PREFIX, n * sizeof(Fq), &pf));
  uint64_t T = (n + 63) / 64;
  uint64_t minT = (uint64_t)ctx->sm_count * 256;
  if (T < minT) T = n < minT ? n : minT;
  k_batch_affine<<<blocks_for(T, 128), 128, 0, ctx->stream>>>(in_dev, n, out_dev, reinterpret_cast<Fq*>(pf), T);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

int32_t compress_points(kzg_ctx* ctx, const G1Affine* in_dev, uint64_t n, uint8_t* out_dev) {
  if (n == 0) return KZG_OK;
  k_compress<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in_dev, n, out_dev);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

int32_t finalize_points(kzg_ctx* ctx, const G1Xyzz* in_dev, uint64_t count, uint8_t* out_c48_dev) {
  if (count == 0) return KZG_OK;
  k_compress_xyzz<<<blocks_for(count, 64), 64, 0, ctx->stream>>>(in_dev, count, out_c48_dev);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

int32_t uncompress_points(kzg_ctx* ctx, const uint8_t* in_dev, uint64_t n, G1Affine* out_dev, int* bad_index) {
  *bad_index = -1;
  if (n == 0) return KZG_OK;
  void* m;
  KZG_TRY(scratch_get(ctx, SCR_MISC, 256, &m));
  unsigned long long init = ~0ull;
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(m, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
  k_uncompress<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(in_dev, n, out_dev, reinterpret_cast<unsigned long long*>(m));
  KZG_LAUNCHED(ctx);
  unsigned long long bad = 0;
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(&bad, m, 8, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad != ~0ull) *bad_index = (int)bad;
  return KZG_OK;
}

// folded-window table into dst[W*n]: dst[j*n + i] = 2^(c j) * src[i]
static int32_t fold_windows(kzg_ctx* ctx, const G1Affine* src, uint64_t n, int c, int W, G1Affine* dst) {
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, n * sizeof(G1Affine), cudaMemcpyDeviceToDevice, ctx->stream));
  if (W <= 1) return KZG_OK;
  void* tmp;
  KZG_TRY(scratch_get(ctx, SCR_XYZZ_TMP, n * sizeof(G1Xyzz), &tmp));
  for (int j = 1; j < W; j++) {
    k_dbl_c<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(dst + (size_t)(j - 1) * n, n, c, reinterpret_cast<G1Xyzz*>(tmp));
    KZG_LAUNCHED(ctx);
    KZG_TRY(batch_to_affine(ctx, reinterpret_cast<G1Xyzz*>(tmp), n, dst + (size_t)j * n));
  }
  return KZG_OK;
}

// multiples table: dst[(b << log2(half)) + d-1] = d * B_b for b in [0, nb), B_b from a folded table
static int32_t fill_multiples(kzg_ctx* ctx, const G1Affine* folded, uint64_t n_stride, int W, uint64_t nb,
                              uint32_t half, G1Affine* dst) {
  uint32_t K = half < 64 ? half : 64;
  uint64_t slab_entries = 1ull << 22;
  uint64_t b_per_slab = slab_entries / half;
  if (b_per_slab == 0) b_per_slab = 1;
  void* tmp;
  KZG_TRY(scratch_get(ctx, SCR_XYZZ_TMP, b_per_slab * half * sizeof(G1Xyzz), &tmp));
  for (uint64_t b0 = 0; b0 < nb; b0 += b_per_slab) {
    uint64_t bc = nb - b0 < b_per_slab ? nb - b0 : b_per_slab;
    uint64_t threads = bc * (half / K);
    k_fill_multiples<<<blocks_for(threads, 128), 128, 0, ctx->stream>>>(folded, n_stride, (uint32_t)W, b0, bc, half, K,
                                                                       reinterpret_cast<G1Xyzz*>(tmp));
    KZG_LAUNCHED(ctx);
    KZG_TRY(batch_to_affine(ctx, reinterpret_cast<G1Xyzz*>(tmp), bc * half, dst + (size_t)b0 * half));
  }
  return KZG_OK;
}

static int32_t ensure_gtable(kzg_ctx* ctx) {
  if (ctx->gtable) return KZG_OK;
  G1Affine* folded = nullptr;
  KZG_CUDA_OK(ctx, cudaMalloc(&folded, sizeof(G1Affine) * GT_W));
  G1Affine* gen = nullptr;
  KZG_CUDA_OK(ctx, cudaMalloc(&gen, sizeof(G1Affine)));
  k_set_generator<<<1, 1, 0, ctx->stream>>>(gen);
  KZG_LAUNCHED(ctx);
  int32_t st = fold_windows(ctx, gen, 1, GT_C, GT_W, folded);
  G1Affine* table = nullptr;
  if (st == KZG_OK) {
    cudaError_t e = cudaMalloc(&table, sizeof(G1Affine) * (size_t)GT_W * GT_HALF);
    if (e != cudaSuccess) { ctx->err = "cudaMalloc(gtable)"; st = KZG_E_NOMEM; }
  }
  if (st == KZG_OK) st = fill_multiples(ctx, folded, 1, GT_W, GT_W, GT_HALF, table);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(folded); cudaFree(gen);
  if (st != KZG_OK) { if (table) cudaFree(table); return st; }
  ctx->gtable = table;
  return KZG_OK;
}

int32_t setup_generate_dev(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n,
                           G1Affine* bases_dev) {
  if (n == 0) return KZG_OK;
  KZG_TRY(ensure_gtable(ctx));
  // s = secret (big-endian) mod r, Montgomery form (Scalar::from_be_bytes, src/scalar.rs:66-73)
  uint8_t le[32];
  for (int i = 0; i < 32; i++) le[i] = secret_be[31 - i];
  Fr raw, s_mont; fr_load_le(raw, le); to_mont(s_mont, raw);
  void* pw; void* xy;
  KZG_TRY(scratch_get(ctx, SCR_FR_A, n * sizeof(Fr), &pw));
  const uint32_t K = 32;
  uint64_t threads = (n + K - 1) / K;
  k_powers<<<blocks_for(threads, 128), 128, 0, ctx->stream>>>(s_mont, first, n, K, reinterpret_cast<Fr*>(pw));
  KZG_LAUNCHED(ctx);
  // large ranges: the 16 table points of every item are summed by affine pairwise trees with shared inversions
  // (batch_tree.cu, 90 Fq products per point); small ranges: one XYZZ walk per thread + batched to-affine (~165)
  if (ctx->opt_setup_tree_min > 0 && n >= (uint64_t)ctx->opt_setup_tree_min)
    return fixed_base_points_tree(ctx, ctx->gtable, GT_C, GT_W, reinterpret_cast<const uint8_t*>(pw), n, bases_dev);
  KZG_TRY(scratch_get(ctx, SCR_XYZZ_TMP, n * sizeof(G1Xyzz), &xy));
  k_setup_points<<<blocks_for(n, 128), 128, 0, ctx->stream>>>(reinterpret_cast<Fr*>(pw), n, ctx->gtable,
                                                              reinterpret_cast<G1Xyzz*>(xy));
  KZG_LAUNCHED(ctx);
  KZG_TRY(batch_to_affine(ctx, reinterpret_cast<G1Xyzz*>(xy), n, bases_dev));
  return KZG_OK;
}

int auto_msm_window(uint64_t n) {
  // Measured on B200 (profiles/sweep_window_r01.txt): with folded windows the accumulate time is flat in c
  // between 14 and 18 bits, the tail grows with the bucket count, and window sizes whose TOP window holds
  // only a few bits funnel n / 2^t entries into a handful of buckets (the heavy path).  Prefer c = 16
  // (16 windows, 15-bit top window) for large n; scale down so that ~2^18 accumulate threads still get
  // >= 32 entries each for small n.
  // With the scalars folded to 254 bits (fr_fold_half) c = 15 needs 17 windows and c = 17 only 15; measured with the
  // fold (profiles/sweep_lanes_r01.txt): at n = 2^20 c = 17 saves 1/16 of the additions but doubles the bucket
  // reduction and ends up level (5.48 vs 5.51 ms); at n = 2^17 c = 15 wins (1.07 vs 1.10 ms).
  // Round 2: the running-sum level 1 of the bucket reduction (msm_tail.cu k_reduce_l1k) made buckets cheap for calls that
  // run for throughput (5.75 instead of 10+ additions per bucket), and with it c = 17 (15 windows, 65536 buckets) wins
  // at n = 2^20: 5.06 -> 4.77 ms per MSM with eight calls in flight, two fused (profiles/sweep_coalesce_r02.txt); c = 18
  // leaves a 3-bit top window (n / 4 entries per bucket: the heavy path) and loses.
  // With pass-through entries free of products and late rounds of under a wave at 2-8 pairs per thread, short runs cost
  // less and c = 17 also wins at 2^18 / 2^19 (2^19: 2.49 -> 2.40 ms fused, 3.71 -> 3.51 alone), c = 16 ties c = 15 at 2^17
  // with a table one window smaller (profiles/sweep_coalesce_r02.txt (g)).
  if (n >= (1ull << 18)) return 17;
  if (n >= (1ull << 17)) return 16;
  int best = 4; double best_cost = 1e300;
  for (int c = 4; c <= 16; c++) {
    int W = windows_for(c);
    int top_bits = 255 - c * (W - 1);
    double entries = (double)W * (double)n;
    double buckets = (double)(1ull << (c - 1));
    double per_thread = entries / (buckets < 262144.0 ? 262144.0 : buckets);
    if (per_thread < 8) per_thread = 8;
    double cost = entries * (1.0 + 2.0 / sqrt(per_thread)) + 8.0 * buckets;
    if (top_bits < c - 4) cost += (double)n;   // skewed top window -> heavy buckets
    cost += 40000.0 * (c > 10 ? (c - 10) / 5.0 : 0.0);  // extra reduction level
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

int32_t build_msm_table(kzg_ctx* ctx, kzg_setup* s, int c) {
  if (c <= 0) c = ctx->opt_msm_window > 0 ? ctx->opt_msm_window : auto_msm_window(s->n);
  if (c < 2 || c > 24) return fail(ctx, KZG_E_BAD_ARG, "window_bits out of range [2,24]");
  if (s->msm_tab && s->msm_c == c) return KZG_OK;
  if (s->msm_tab) { cudaFree(s->msm_tab); s->msm_tab = nullptr; }
  int W = windows_for(c);
  if (s->n == 0) { s->msm_c = c; s->msm_w = W; return KZG_OK; }
  const uint64_t entries = (uint64_t)W * s->n;
  G1Affine* dense = nullptr;     // built densely (each window is 2^c times the previous one), then spread to the pitch
  char* tab = nullptr;
  if (cudaMalloc(&dense, sizeof(G1Affine) * entries) != cudaSuccess ||
      cudaMalloc(&tab, (size_t)MSM_TAB_PITCH * entries) != cudaSuccess) {
    cudaGetLastError();
    if (dense) cudaFree(dense);
    return fail(ctx, KZG_E_NOMEM, "cudaMalloc(msm table)");
  }
  int32_t st = fold_windows(ctx, s->bases, s->n, c, W, dense);
  if (st == KZG_OK) {
    k_pitch_table<<<blocks_for(entries * 8, 256), 256, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(dense), entries,
                                                                        reinterpret_cast<uint4*>(tab));
    ctx->launches++;
    if (cudaGetLastError() != cudaSuccess) st = fail(ctx, KZG_E_CUDA, "k_pitch_table launch");
  }
  cudaError_t e = cudaStreamSynchronize(ctx->stream);
  cudaFree(dense);
  if (st == KZG_OK && e != cudaSuccess) st = fail(ctx, KZG_E_CUDA, "msm table build");
  if (st != KZG_OK) { cudaFree(tab); return st; }
  s->msm_tab = tab; s->msm_c = c; s->msm_w = W;
  return KZG_OK;
}

// bytes of the fixed-base table for n_points bases at window c (W windows x 2^(c-1) multiples per base and window)
uint64_t fixed_table_bytes(uint64_t n_points, int c) {
  return (uint64_t)windows_for(c) * n_points * ((uint64_t)1 << (c - 1)) * sizeof(G1Affine);
}

// Window of the fixed-base table when the caller does not name one: blob throughput is ~ 1 / W(c) and random
// 96-byte reads cost the same over 100 GiB as over 1 GiB, so the widest window whose table (plus the
// transient folded copy) fits the memory budget wins.  Budget = option "fixed_table_max_bytes", else 66 % of the
// memory that is free right now (cudaMemGetInfo): the rest stays for the batch scratch (~8 GiB at 2048 rows of 4096)
// and for other tenants of the device.  n = 4096 on an empty B200 (178 GiB free): c = 15 (102 GiB); with 100 GiB free:
// c = 14 (57 GiB); with 50 GiB free: c = 13 (30 GiB).
int auto_fixed_window(kzg_ctx* ctx, uint64_t n_points) {
  size_t free_b = 0, total_b = 0;
  uint64_t budget = ctx->opt_fixed_table_max_bytes > 0 ? (uint64_t)ctx->opt_fixed_table_max_bytes : 0;
  if (!budget) {
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) { cudaGetLastError(); return 10; }
    budget = (uint64_t)((double)free_b * 0.66);
  }
  int best = 2;
  for (int c = 2; c <= 15; c++)
    if (fixed_table_bytes(n_points ? n_points : 1, c) <= budget) best = c;
  return best;
}

int32_t build_fixed_table(kzg_ctx* ctx, kzg_setup* s, uint64_t n_points, int c) {
  if (n_points > s->n) n_points = s->n;
  if (c <= 0) c = ctx->opt_fixed_window > 0 ? ctx->opt_fixed_window : auto_fixed_window(ctx, n_points);
  if (c < 2 || c > 16) return fail(ctx, KZG_E_BAD_ARG, "fixed window_bits out of range [2,16]");
  if (s->fix_tab && s->fix_c == c && s->fix_n >= n_points) return KZG_OK;
  if (s->fix_tab) { cudaFree(s->fix_tab); s->fix_tab = nullptr; s->fix_n = 0; }
  if (n_points == 0) return KZG_OK;
  int W = windows_for(c);
  uint32_t half = 1u << (c - 1);
  G1Affine* folded = nullptr;
  G1Affine* tab = nullptr;
  if (cudaMalloc(&folded, sizeof(G1Affine) * (size_t)W * n_points) != cudaSuccess) {
    cudaGetLastError(); return fail(ctx, KZG_E_NOMEM, "cudaMalloc(fixed fold)");
  }
  if (cudaMalloc(&tab, sizeof(G1Affine) * (size_t)W * n_points * half) != cudaSuccess) {
    cudaGetLastError(); cudaFree(folded); return fail(ctx, KZG_E_NOMEM, "cudaMalloc(fixed table)");
  }
  int32_t st = fold_windows(ctx, s->bases, n_points, c, W, folded);
  if (st == KZG_OK) st = fill_multiples(ctx, folded, n_points, W, n_points * (uint64_t)W, half, tab);
  cudaStreamSynchronize(ctx->stream);
  cudaFree(folded);
  if (st != KZG_OK) { cudaFree(tab); return st; }
  s->fix_tab = tab; s->fix_c = c; s->fix_w = W; s->fix_n = n_points;
  return KZG_OK;
}

}  // namespace kzg
