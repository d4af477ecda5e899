// K2' — the first rounds of the bucket accumulation in affine coordinates.
//
// After the radix sort every bucket owns a contiguous run of (table index, sign) entries; all buckets together
// hold W*n of them (16.8 M at n = 2^20).  k_bucket_acc adds them one by one into XYZZ registers: 10 Fq products
// per addition.  Here the runs are first HALVED a few times: neighbouring entries of a run are added pairwise in
// affine coordinates, all pairs of all buckets in one grid, with one shared inversion per round (Montgomery's
// trick across the grid, batch_tree.cu) — 6.6 products per addition.  Three rounds do 7/8 of all additions;
// what is left (W*n/8 dense affine points, same bucket structure) goes to k_bucket_acc and the usual tail.
// Only worth it when a round has millions of independent additions (its inversion chain is ~0.15 ms of
// latency): msm_bucket runs one round per doubling of the entry count from 2^21 entries on (4 rounds at n = 2^20:
// -18 %; break-even at 2^21).  Pairs never cross a bucket; a run of odd length passes
// its last entry through; additions are complete (infinity, P = Q, P = -Q).
#include "msm_tree_kernels.cuh"

namespace kzg {

// bucket bounds of every round at once: with len_0(b) = seg0[b+1] - seg0[b] and len_r = ceil(len_{r-1} / 2),
// segs[r][b] = sum_{b' < b} len_{r+1}(b'), b = 0 .. M, r < R.  Three small launches (a single-block version took
// 190 us of pure load latency): per-tile totals, scan of the tile totals, per-tile scans + offsets.
constexpr int SEG_TILE = 1024;
constexpr int SEG_MAX_ROUNDS = 16;

__device__ __forceinline__ uint32_t halved(uint32_t len, int times) {
  for (int q = 0; q < times; q++) len = (len + 1) >> 1;
  return len;
}
// block-wide exclusive scan of one value per thread (1024 threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* sh /* 33 words */) {
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += t; }
  if (lane == 31) sh[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = sh[lane], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= (uint32_t)d) xi += t; }
    sh[lane] = xi - x;                 // exclusive prefix of the warp totals
    if (lane == 31) sh[32] = xi;
  }
  __syncthreads();
  uint32_t res = sh[w] + inc - v;
  *total = sh[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SEG_TILE)
k_seg_tile_sums(const uint32_t* __restrict__ seg0, uint32_t M, int R, uint32_t ntiles, uint32_t* __restrict__ tile_tot) {
  __shared__ uint32_t sh[33];
  const uint32_t b = blockIdx.x * SEG_TILE + threadIdx.x;
  const uint32_t len0 = b < M ? __ldg(seg0 + b + 1) - __ldg(seg0 + b) : 0;
  for (int r = 0; r < R; r++) {
    uint32_t tot;
    block_excl_scan(halved(len0, r + 1), &tot, sh);
    if (threadIdx.x == 0) tile_tot[(size_t)r * ntiles + blockIdx.x] = tot;
  }
}
// one block: exclusive scan of the tile totals of every round (ntiles <= 1024), in place
__global__ void __launch_bounds__(SEG_TILE)
k_seg_tile_scan(uint32_t* __restrict__ tile_tot, uint32_t ntiles, int R, uint32_t M, uint32_t* __restrict__ segs) {
  __shared__ uint32_t sh[33];
  for (int r = 0; r < R; r++) {
    uint32_t v = threadIdx.x < ntiles ? tile_tot[(size_t)r * ntiles + threadIdx.x] : 0;
    uint32_t tot;
    uint32_t ex = block_excl_scan(v, &tot, sh);
    if (threadIdx.x < ntiles) tile_tot[(size_t)r * ntiles + threadIdx.x] = ex;
    if (threadIdx.x == 0) segs[(size_t)r * (M + 2) + M] = tot;
  }
}
__global__ void __launch_bounds__(SEG_TILE)
k_seg_write(const uint32_t* __restrict__ seg0, uint32_t M, int R, uint32_t ntiles, const uint32_t* __restrict__ tile_off,
            uint32_t* __restrict__ segs) {
  __shared__ uint32_t sh[33];
  const uint32_t b = blockIdx.x * SEG_TILE + threadIdx.x;
  const uint32_t len0 = b < M ? __ldg(seg0 + b + 1) - __ldg(seg0 + b) : 0;
  for (int r = 0; r < R; r++) {
    uint32_t tot;
    uint32_t ex = block_excl_scan(halved(len0, r + 1), &tot, sh);
    if (b < M) segs[(size_t)r * (M + 2) + b] = ex + tile_off[(size_t)r * ntiles + blockIdx.x];
  }
}

__global__ void k_iota(uint32_t* out, uint64_t n) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (uint32_t)i;
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

// msm_tree_r1.cu: k_seg_add_lean<true, 4> compiled with the field product out of line (option tree_r1_ool)
void launch_seg_add_r1_ool(cudaStream_t stream, unsigned grid, const uint32_t* seg_in, const uint32_t* seg_out, uint32_t M,
                           const uint32_t* vals, const G1Affine* src, uint64_t T, int K, const Fq* prefix, Fq* tot,
                           G1Affine* out, uint32_t pitch);

// Which addition kernel a round runs (options tree_lean / tree_lean_r1; measured at n = 2^20, profiles/sweep_lean_r02.txt):
// the reloading variant at the SAME 4 blocks per SM is 0.5-0.8 % faster on the dense rounds (6.46 -> 6.40 ms alone,
// 5.40 -> 5.36 ms with four lanes; on round 1 another 6.39 -> 6.31 / 5.37 -> 5.35 ms) and is the default; at 5 or 6 blocks
// per SM it is SLOWER (dense rounds 5.40 / 5.46 ms, round 1 5.53 / 5.64 ms): more warps do not help, the multiplier pipe is the bound, not latency.  Capping the
// residency of these kernels with idle shared memory so that another lane's denominator gather could run beside them
// costs 6 % (profiles/sweep_carve_r02.txt): not kept.
template <bool R1>
static void launch_seg_add(kzg_ctx* ctx, unsigned grid, const uint32_t* seg_in, const uint32_t* so, uint32_t M, const uint32_t* vals,
                           const G1Affine* src, uint64_t T, int K, const Fq* prefix, Fq* tot, G1Affine* out, uint32_t pitch) {
  const int lean = R1 ? ctx->opt_tree_lean_r1 : ctx->opt_tree_lean;
  if (R1 && lean == 4 && ctx->opt_tree_r1_ool) { launch_seg_add_r1_ool(ctx->stream, grid, seg_in, so, M, vals, src, T, K, prefix, tot, out, pitch); return; }
  if (lean == 6) k_seg_add_lean<R1, 6><<<grid, 128, 0, ctx->stream>>>(seg_in, so, M, vals, src, T, K, prefix, tot, out, pitch);
  else if (lean == 5) k_seg_add_lean<R1, 5><<<grid, 128, 0, ctx->stream>>>(seg_in, so, M, vals, src, T, K, prefix, tot, out, pitch);
  else if (lean == 4) k_seg_add_lean<R1, 4><<<grid, 128, 0, ctx->stream>>>(seg_in, so, M, vals, src, T, K, prefix, tot, out, pitch);
  else k_seg_add<R1><<<grid, 128, 0, ctx->stream>>>(seg_in, so, M, vals, src, T, K, prefix, tot, out, pitch);
}

int32_t msm_tree_rounds(kzg_ctx* ctx, const uint32_t* start_dev, const uint32_t* vals_dev, const char* tab_pitched,
                        uint32_t M, uint64_t entries, int rounds, const G1Affine** pts_out,
                        const uint32_t** seg_out_p, const uint32_t** iota_out, uint64_t* entries_out,
                        uint32_t tab_pitch, int level) {
  // level 1: the scratch arenas SCR_TREE_*; level 2 (the bucket sums of level 1 reduced as a bucket problem of their own,
  // msm.cu finish_level2) reads level 1's arenas while it runs, so it has a set of its own
  const int sA = level == 2 ? SCR_L2_A : SCR_TREE_A, sB = level == 2 ? SCR_L2_B : SCR_TREE_B;
  const int sPF = level == 2 ? SCR_L2_PREFIX : SCR_TREE_PREFIX, sTW = level == 2 ? SCR_L2_TOT : SCR_TREE_TOT;
  const int sSG = level == 2 ? SCR_L2_SEG : SCR_TREE_SEG, sIO = level == 2 ? SCR_L2_IOTA : SCR_TREE_IOTA;
  const uint64_t top = (uint64_t)ctx->opt_inv_top;
  // upper bounds on the number of points after each round (exact counts stay on the device)
  auto halve = [M](uint64_t e) { uint64_t h = (e + M) / 2 + 1; return h < e ? h : e; };   // never grows
  const uint64_t e1 = halve(entries);
  const uint64_t e2 = halve(e1);
  const uint64_t T1 = (e1 + 8 - 1) / 8;            // most threads a round can have (K >= 8)
  void *A, *B, *PF, *TW, *SG, *IO;
  KZG_TRY(scratch_get(ctx, sA, e1 * sizeof(G1Affine), &A));
  KZG_TRY(scratch_get(ctx, sB, e2 * sizeof(G1Affine), &B));
  KZG_TRY(scratch_get(ctx, sPF, (T1 * ST_K + 16) * sizeof(Fq), &PF));
  KZG_TRY(scratch_get(ctx, sTW, (T1 + invert_scratch_elems(T1, top)) * sizeof(Fq), &TW));
  if (rounds > SEG_MAX_ROUNDS) rounds = SEG_MAX_ROUNDS;
  const uint32_t ntiles = (M + SEG_TILE - 1) / SEG_TILE;
  if (ntiles > SEG_TILE) return fail(ctx, KZG_E_BAD_ARG, "too many buckets for the halving rounds");
  KZG_TRY(scratch_get(ctx, sSG, (size_t)rounds * ((size_t)M + 2 + ntiles) * 4, &SG));
  Fq* prefix = reinterpret_cast<Fq*>(PF);
  Fq* tot = reinterpret_cast<Fq*>(TW);
  Fq* inv_ws = tot + T1;
  uint32_t* segs = reinterpret_cast<uint32_t*>(SG);
  uint32_t* tile_tot = segs + (size_t)rounds * (M + 2);
  k_seg_tile_sums<<<ntiles, SEG_TILE, 0, ctx->stream>>>(start_dev, M, rounds, ntiles, tile_tot);
  KZG_LAUNCHED(ctx);
  k_seg_tile_scan<<<1, SEG_TILE, 0, ctx->stream>>>(tile_tot, ntiles, rounds, M, segs);
  KZG_LAUNCHED(ctx);
  k_seg_write<<<ntiles, SEG_TILE, 0, ctx->stream>>>(start_dev, M, rounds, ntiles, tile_tot, segs);
  KZG_LAUNCHED(ctx);
  G1Affine* buf[2] = {reinterpret_cast<G1Affine*>(A), reinterpret_cast<G1Affine*>(B)};

  const uint32_t* seg_in = start_dev;
  const G1Affine* src = reinterpret_cast<const G1Affine*>(tab_pitched);   // round 1 indexes it by MSM_TAB_PITCH
  uint64_t bound = entries;
  for (int r = 0; r < rounds; r++) {
    bound = halve(bound);
    // pairs per thread: the blocks of a launch are identical and run in lockstep, so a launch costs WHOLE waves of
    // (4 - 6 blocks per SM) x K pair-times — 3.47 waves cost 4.  Pick the K <= 16 that fills the last wave.
    int K = ST_K;
    {
      const int lean = r == 0 ? ctx->opt_tree_lean_r1 : ctx->opt_tree_lean;
      const int bps = lean >= 4 && lean <= 6 ? lean : 4;                 // resident blocks per SM of this round's addition kernel
      const double per_slot = (double)bound * 1.004 / ((double)ctx->sm_count * bps * 128);
      int waves = (int)((per_slot + ST_K - 1e-9) / ST_K); if (waves < 1) waves = 1;
      K = (int)(per_slot / waves + 0.999999);
      // at least 8 pairs per thread in round 1 (the arenas hold e1 / 8 thread totals); a late round of under a wave may
      // go down to 2: its threads then run 2 instead of 8 additions back to back (a lone warp needs ~7 us per addition)
      int kmin = (int)((8.0 * (double)bound) / (double)e1 + 0.999999);
      if (kmin < 2) kmin = 2;
      if (kmin > 8) kmin = 8;
      if (K < kmin) K = kmin;
      if (K > ST_K) K = ST_K;
      if (ctx->opt_tree_k >= 8 && ctx->opt_tree_k <= ST_K) K = ctx->opt_tree_k;
    }
    const uint64_t T = (bound + K - 1) / K;
    uint32_t* so = segs + (size_t)r * (M + 2);
    G1Affine* out = buf[r & 1];
    if (r == 0 && level == 1 && ctx->opt_den_side > 0 && ctx->throughput_mode && blocks_for(T, 128) > 4u * (unsigned)ctx->sm_count) {
      // The round-1 denominators are a gather at the DRAM random-access rate with one product per pair: the multiplier idles.
      // With calls in flight, run them as a few resident blocks per SM on a HIGH-priority side stream, so that they sit
      // beside another lane's addition kernel (which keeps the other block slots) instead of owning the machine.
      if (!ctx->side_stream) {
        int lo = 0, hi = 0;
        KZG_CUDA_OK(ctx, cudaDeviceGetStreamPriorityRange(&lo, &hi));
        KZG_CUDA_OK(ctx, cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, hi));
        KZG_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->side_ev[0], cudaEventDisableTiming));
        KZG_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->side_ev[1], cudaEventDisableTiming));
      }
      KZG_CUDA_OK(ctx, cudaEventRecord(ctx->side_ev[0], ctx->stream));
      KZG_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->side_ev[0], 0));
      k_seg_den<true><<<(unsigned)(ctx->sm_count * ctx->opt_den_side), 128, 0, ctx->side_stream>>>(seg_in, so, M, vals_dev, src, T, K, prefix, tot, tab_pitch);
      KZG_CUDA_OK(ctx, cudaEventRecord(ctx->side_ev[1], ctx->side_stream));
      KZG_CUDA_OK(ctx, cudaStreamWaitEvent(ctx->stream, ctx->side_ev[1], 0));
    }
    else if (r == 0) k_seg_den<true><<<blocks_for(T, 128), 128, 0, ctx->stream>>>(seg_in, so, M, vals_dev, src, T, K, prefix, tot, tab_pitch);
    else k_seg_den<false><<<blocks_for(T, 128), 128, 0, ctx->stream>>>(seg_in, so, M, nullptr, src, T, K, prefix, tot, 0);
    KZG_LAUNCHED(ctx);
    KZG_TRY(batch_invert_inplace(ctx, tot, T, inv_ws, top));
    if (r == 0) {
      if (level == 1) KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[0], ctx->stream));
      launch_seg_add<true>(ctx, blocks_for(T, 128), seg_in, so, M, vals_dev, src, T, K, prefix, tot, out, tab_pitch);
      if (level == 1) KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[1], ctx->stream));
    } else launch_seg_add<false>(ctx, blocks_for(T, 128), seg_in, so, M, nullptr, src, T, K, prefix, tot, out, 0);
    KZG_LAUNCHED(ctx);
    seg_in = so; src = out;
  }
  KZG_TRY(scratch_get(ctx, sIO, (bound + 16) * 4, &IO));
  k_iota<<<blocks_for(bound, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(IO), bound);
  KZG_LAUNCHED(ctx);
  *pts_out = src; *seg_out_p = seg_in; *iota_out = reinterpret_cast<uint32_t*>(IO); *entries_out = bound;
  return KZG_OK;
}

}  // namespace kzg
