// K2 — Pippenger G1 multi-scalar multiplication (the commit / open hot loop).
//
// Replaces Polynomial::commit's loop (src/polynomial.rs:207-212): n x (256-bit blst_p1_mult +
// blst_p1_add_or_double), i.e. O(255 n) group operations, by
//   1. recode   : canonical scalar, folded to min(k, r-k) -> W = ceil(255/c) signed c-bit digits; digit d of
//                 window j for point i becomes the pair (bucket |d|-1, table index j*n+i | sign).
//                 Windows are FOLDED: the setup holds 2^(c j) * P_i for every j (built once per
//                 setup, setup.cu), so all W windows share one set of 2^(c-1) buckets and the
//                 per-window reduction + c doublings per window disappear.
//   2. sort     : counting sort of the (bucket, index) entries (count / tiled scan / scatter; cub::DeviceRadixSort behind
//                 option msm_sort = 1).
//   3. bounds   : bucket b owns sorted entries [start[b], start[b+1]).
//   4. accumulate: one thread per bucket, XYZZ += affine (8M+2S), table points fetched with
//                 128-bit loads; buckets above `heavy_cap` entries (adversarial inputs: all-equal
//                 scalars) go to a block-per-bucket kernel with a shared-memory tree.
//   5. reduce   : sum_b (b+1) * B_b by warp-level suffix scans, three fused launches (msm_tail.cu).
// Integer-multiply bound: 3 000 wide multiply-adds per bucket addition, 96 B of HBM per addition.
#include <cub/device/device_radix_sort.cuh>
#include <algorithm>
#include "engine.cuh"

namespace kzg {

__device__ __forceinline__ void ld_affine_g(G1Affine& o, const G1Affine* p) {
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

__device__ __forceinline__ void ld_scalar(Fr& o, const uint8_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
}

// ---- 1. recode -------------------------------------------------------------------------------
__global__ void k_recode(const uint8_t* __restrict__ scalars, uint64_t n, int c, int W, uint64_t n_stride,
                         uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k; ld_scalar(k, scalars + i * 32);
  fr_canon_reduce(k);
  SignedDigits sd; sd.init(k, c);                          // k <- min(k, r - k): one bit less to recode
  const uint32_t half = 1u << (c - 1);
  for (int j = 0; j < W; j++) {
    bool neg;
    const uint32_t d = sd.next(j, &neg);
    keys[(size_t)j * n + i] = d ? d - 1 : half;            // `half` = discard bucket for zero digits
    vals[(size_t)j * n + i] = (uint32_t)((uint64_t)j * n_stride + i) | (neg ? 0x80000000u : 0u);
  }
}

// ---- 3. bucket bounds --------------------------------------------------------------------------
__global__ void k_bucket_bounds(const uint32_t* __restrict__ keys, uint64_t total, uint32_t nb,
                                uint32_t* __restrict__ start) {
  uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  uint32_t k = keys[e];
  uint32_t lo = e ? keys[e - 1] + 1 : 0;
  for (uint32_t b = lo; b <= k; b++) start[b] = (uint32_t)e;
  if (e == total - 1)
    for (uint32_t b = k + 1; b <= nb; b++) start[b] = (uint32_t)total;
}

// ---- 1-3 in one: counting sort ------------------------------------------------------------------
// The keys are bucket numbers, at most 2^16 distinct values, so no radix passes are needed: count, scan, scatter.
//   k_digit_count   recodes every scalar and adds one to count[bucket] per non-zero digit (global RED, no return value);
//   k_bucket_tile_totals + k_bucket_scan   exclusive scan of the counts over tiles of 1024 -> start[0..M] and the scatter cursors;
//   k_digit_scatter recodes again (32 B of scalar re-read per point instead of 64 B of digits written and read back) and
//                   reserves each entry's slot in its bucket's run with one ATOM: vals[slot] = table index | sign.
// Zero digits simply vanish (the radix path parks them in a discard bucket).  Entries of a bucket arrive in whatever
// order the atomics resolve: sums do not care, every addition downstream is complete, and the result is a canonical
// encoding — bit-identical output, fewer passes: recode + 2 onesweep passes + histogram + bounds (0.37 ms at n = 2^20)
// become three launches.
// Several calls at once (msm_bucket_multi): blockIdx.y = call, its buckets are [call * half, (call + 1) * half).
struct ScalarPtrs { const uint8_t* p[KZG_MAX_COALESCE]; };

__global__ void __launch_bounds__(256)
k_digit_count(ScalarPtrs sp, uint64_t n, int c, int W, uint32_t half, uint32_t* __restrict__ count) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* __restrict__ scalars = sp.p[blockIdx.y];
  count += (size_t)blockIdx.y * half;
  Fr k; ld_scalar(k, scalars + i * 32);
  fr_canon_reduce(k);
  SignedDigits sd; sd.init(k, c);
#pragma unroll 1
  for (int j = 0; j < W; j++) {
    bool neg;
    const uint32_t d = sd.next(j, &neg);
    if (d) atomicAdd(count + (d - 1), 1u);
  }
}

// start[b] = sum of count[< b], start[M] = all entries; cursor = start.  Two launches over tiles of 1024 counts (a single
// block took 164 us of load latency at M = 65536, the buckets of four fused calls): per-tile totals, then every tile adds
// up the totals in front of it and scans its own counts.
constexpr int SCAN_TILE = 1024;
__global__ void __launch_bounds__(SCAN_TILE)
k_bucket_tile_totals(const uint32_t* __restrict__ count, uint32_t M, uint32_t* __restrict__ tile_tot) {
  __shared__ uint32_t warp_tot[32];
  const uint32_t b = blockIdx.x * SCAN_TILE + threadIdx.x;
  uint32_t v = b < M ? __ldg(count + b) : 0u;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if ((threadIdx.x & 31) == 0) warp_tot[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = warp_tot[threadIdx.x];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (threadIdx.x == 0) tile_tot[blockIdx.x] = v;
  }
}
// ntiles <= 1024
__global__ void __launch_bounds__(SCAN_TILE)
k_bucket_scan(const uint32_t* __restrict__ count, uint32_t M, const uint32_t* __restrict__ tile_tot,
              uint32_t* __restrict__ start, uint32_t* __restrict__ cursor) {
  __shared__ uint32_t warp_tot[32];
  __shared__ uint32_t base_sh;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  // totals of the tiles in front of this one
  uint32_t v = threadIdx.x < blockIdx.x ? __ldg(tile_tot + threadIdx.x) : 0u;
#pragma unroll
  for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
  if (lane == 0) warp_tot[w] = v;
  __syncthreads();
  if (w == 0) {
    v = warp_tot[lane];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if (lane == 0) base_sh = v;
  }
  __syncthreads();
  const uint32_t base = base_sh;
  const uint32_t b = blockIdx.x * SCAN_TILE + threadIdx.x;
  const uint32_t local = b < M ? __ldg(count + b) : 0u;
  uint32_t inc = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += t; }
  __syncthreads();
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = warp_tot[lane], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= (uint32_t)d) xi += t; }
    warp_tot[lane] = xi - x;
  }
  __syncthreads();
  const uint32_t ex = base + warp_tot[w] + inc - local;
  if (b < M) { start[b] = ex; cursor[b] = ex; }
  if (b == M - 1) start[M] = ex + local;
}

__global__ void __launch_bounds__(256)
k_digit_scatter(ScalarPtrs sp, uint64_t n, int c, int W, uint64_t n_stride, uint32_t half,
                uint32_t* __restrict__ cursor, uint32_t* __restrict__ vals) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* __restrict__ scalars = sp.p[blockIdx.y];
  cursor += (size_t)blockIdx.y * half;
  Fr k; ld_scalar(k, scalars + i * 32);
  fr_canon_reduce(k);
  SignedDigits sd; sd.init(k, c);
  // four digits at a time: their atomics are in flight together (the returned slot is only needed for the store)
#pragma unroll 1
  for (int j0 = 0; j0 < W; j0 += 4) {
    uint32_t dd[4], vv[4], slot[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      dd[u] = 0; vv[u] = 0;
      if (j0 + u < W) {
        bool neg;
        dd[u] = sd.next(j0 + u, &neg);
        vv[u] = (uint32_t)((uint64_t)(j0 + u) * n_stride + i) | (neg ? 0x80000000u : 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; u++) if (dd[u]) slot[u] = atomicAdd(cursor + (dd[u] - 1), 1u);
#pragma unroll
    for (int u = 0; u < 4; u++) if (dd[u]) vals[slot[u]] = vv[u];
  }
}

// ---- 1-3 in one, two levels: partition sort ----------------------------------------------------------------------
// The counting sort above pays two GLOBAL atomics per entry (a RED to count, an ATOM to reserve the slot): 21 ps per entry,
// bound by the L2 atomic units.  Here the buckets are first split into PARTITIONS of 128 consecutive buckets:
//   k_part_count    per block of 256 scalars a shared-memory histogram over the partitions, then one global RED per
//                   (block, non-empty partition) instead of one per entry;
//   k_part_scan     one block: partition bounds;
//   k_part_scatter  the same histogram again, one global ATOM per (block, partition) reserves the block's range in the
//                   partition, shared-memory atomics hand out the slots inside it: (bucket mod 128, table index | sign) as
//                   one 8-byte record per entry, written in runs of ~15 per partition and block;
//   k_part_sort     one block per partition (~60 k records): histogram of its 512 buckets in shared memory -> their
//                   bounds (start[]) -> records scattered into the final entry order, all atomics in shared memory.
// Same contract as the counting sort: start[0..M], vals in bucket order (any order inside a bucket), zero digits vanish.
// MEASURED (profiles/sweep_coalesce_r02.txt (h), two fused calls at n = 2^20, 31.5 M entries): k_part_count 70 us (the
// counting sort's count: 214) and k_part_sort 116, but k_part_scatter 356 (two recodes, two shared atomics and an 8-byte
// store per entry at one 1024-thread block per SM) — 545 us against 672 for the counting sort, which four lanes do not
// see (4.68 vs 4.64 ms per MSM: the 192 KB blocks of k_part_sort share an SM with nobody); a lone call gains 0.05 ms.
// Kept behind option msm_sort = 2.  (First version, 512-bucket partitions without staging: 65 + 352 + 304 us.)
constexpr int PART_LOG = 7;                     // buckets per partition (log2)
constexpr uint32_t PART_W = 1u << PART_LOG;
constexpr int PART_MAX = 8192;                  // partitions (M <= 2^20 buckets)
constexpr int PART_TILE = 1024;                 // scalars per block of the two passes over the scalars
constexpr uint32_t PART_STAGE = 48 * 1024;      // records of a partition whose final order is staged in shared memory (192 KB)

__global__ void __launch_bounds__(PART_TILE)
k_part_count(ScalarPtrs sp, uint64_t n, int c, int W, uint32_t half, uint32_t* __restrict__ part_cnt) {
  __shared__ uint32_t hist[PART_MAX];
  const uint32_t np_call = half >> PART_LOG;                       // partitions of one call
  for (uint32_t k = threadIdx.x; k < np_call; k += PART_TILE) hist[k] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    Fr k; ld_scalar(k, sp.p[blockIdx.y] + i * 32);
    fr_canon_reduce(k);
    SignedDigits sd; sd.init(k, c);
#pragma unroll 1
    for (int j = 0; j < W; j++) {
      bool neg;
      const uint32_t d = sd.next(j, &neg);
      if (d) atomicAdd(&hist[(d - 1) >> PART_LOG], 1u);
    }
  }
  __syncthreads();
  part_cnt += (size_t)blockIdx.y * np_call;
  for (uint32_t k = threadIdx.x; k < np_call; k += PART_TILE) if (hist[k]) atomicAdd(part_cnt + k, hist[k]);
}

// NP <= PART_MAX counts -> part_start[0..NP] and the scatter cursors; one block of 1024 threads, consecutive counts per thread
__global__ void __launch_bounds__(1024)
k_part_scan(const uint32_t* __restrict__ part_cnt, uint32_t NP, uint32_t* __restrict__ part_start, uint32_t* __restrict__ part_cur) {
  __shared__ uint32_t warp_tot[32];
  const uint32_t per = (NP + 1023) / 1024, b0 = threadIdx.x * per;
  uint32_t local = 0;
  for (uint32_t k = 0; k < per; k++) if (b0 + k < NP) local += part_cnt[b0 + k];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = local;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += t; }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = warp_tot[lane], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= (uint32_t)d) xi += t; }
    warp_tot[lane] = xi - x;
  }
  __syncthreads();
  uint32_t run = warp_tot[w] + inc - local;
  for (uint32_t k = 0; k < per; k++) if (b0 + k < NP) { part_start[b0 + k] = run; part_cur[b0 + k] = run; run += part_cnt[b0 + k]; }
  if (threadIdx.x == 1023) part_start[NP] = run;                  // the last thread's running sum is the grand total
}

__global__ void __launch_bounds__(PART_TILE)
k_part_scatter(ScalarPtrs sp, uint64_t n, int c, int W, uint64_t n_stride, uint32_t half,
               uint32_t* __restrict__ part_cur, uint2* __restrict__ rec) {
  __shared__ uint32_t hist[PART_MAX];      // counts, then the block's base inside each partition, then the running slot
  const uint32_t np_call = half >> PART_LOG;
  for (uint32_t k = threadIdx.x; k < np_call; k += PART_TILE) hist[k] = 0;
  __syncthreads();
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fr k;
  if (i < n) {
    ld_scalar(k, sp.p[blockIdx.y] + i * 32);
    fr_canon_reduce(k);
    SignedDigits sd; sd.init(k, c);
#pragma unroll 1
    for (int j = 0; j < W; j++) {
      bool neg;
      const uint32_t d = sd.next(j, &neg);
      if (d) atomicAdd(&hist[(d - 1) >> PART_LOG], 1u);
    }
  }
  __syncthreads();
  part_cur += (size_t)blockIdx.y * np_call;
  for (uint32_t q = threadIdx.x; q < np_call; q += PART_TILE) {
    const uint32_t h = hist[q];
    hist[q] = h ? atomicAdd(part_cur + q, h) : 0u;               // this block's range in partition q starts here
  }
  __syncthreads();
  if (i < n) {
    SignedDigits sd; sd.init(k, c);
#pragma unroll 1
    for (int j = 0; j < W; j++) {
      bool neg;
      const uint32_t d = sd.next(j, &neg);
      if (!d) continue;
      const uint32_t slot = atomicAdd(&hist[(d - 1) >> PART_LOG], 1u);
      rec[slot] = make_uint2((uint32_t)((uint64_t)j * n_stride + i) | (neg ? 0x80000000u : 0u), (d - 1) & (PART_W - 1u));
    }
  }
}

// one block per partition: records [part_start[P], part_start[P+1]) -> start[P * 128 .. ] and vals in bucket order.  A
// partition of up to PART_STAGE records gets its final order in shared memory and leaves in one coalesced copy (random
// 4-byte stores to global memory cost a 32-byte sector each); a larger one (skewed scalars) is scattered directly.
__global__ void __launch_bounds__(1024)
k_part_sort(const uint2* __restrict__ rec, const uint32_t* __restrict__ part_start, uint32_t NP,
            uint32_t* __restrict__ start, uint32_t* __restrict__ vals) {
  extern __shared__ uint32_t stage[];           // PART_STAGE words
  __shared__ uint32_t cnt[PART_W];
  __shared__ uint32_t warp_tot[32];
  const uint32_t P = blockIdx.x;
  const uint32_t lo = __ldg(part_start + P), hi = __ldg(part_start + P + 1);
  if (threadIdx.x < PART_W) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (uint32_t e = lo + threadIdx.x; e < hi; e += 1024) atomicAdd(&cnt[rec[e].y], 1u);
  __syncthreads();
  // exclusive scan of the 128 counts by the first four warps
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t own = threadIdx.x < PART_W ? cnt[threadIdx.x] : 0u, inc = own;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += t; }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = warp_tot[lane], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= (uint32_t)d) xi += t; }
    warp_tot[lane] = xi - x;
  }
  __syncthreads();
  const bool staged = hi - lo <= PART_STAGE;
  if (threadIdx.x < PART_W) {
    const uint32_t ex = warp_tot[w] + inc - own;                 // position inside the partition
    start[(size_t)P * PART_W + threadIdx.x] = lo + ex;
    cnt[threadIdx.x] = staged ? ex : lo + ex;                    // running slot of the bucket
  }
  if (P == NP - 1 && threadIdx.x == 0) start[(size_t)NP * PART_W] = hi;
  __syncthreads();
  if (staged) {
    for (uint32_t e = lo + threadIdx.x; e < hi; e += 1024) {
      const uint2 r = rec[e];
      stage[atomicAdd(&cnt[r.y], 1u)] = r.x;
    }
    __syncthreads();
    for (uint32_t e = threadIdx.x; e < hi - lo; e += 1024) vals[lo + e] = stage[e];
  } else {
    for (uint32_t e = lo + threadIdx.x; e < hi; e += 1024) {
      const uint2 r = rec[e];
      vals[atomicAdd(&cnt[r.y], 1u)] = r.x;
    }
  }
}

static inline unsigned blocks_for(uint64_t n, unsigned bs);
// buckets of the calls sp.p[0..B) (half each) -> start[0 .. B * half], vals; false if the geometry does not fit (the caller
// then takes the counting sort)
static bool part_sort_fits(uint32_t half, int B) {
  const uint64_t M = (uint64_t)half * (uint64_t)B;
  return half >= 8 * PART_W && (M >> PART_LOG) <= (uint64_t)PART_MAX;
}
static int32_t launch_part_sort(kzg_ctx* ctx, const ScalarPtrs& sp, int B, uint64_t n, int c, int W, uint64_t n_stride,
                                uint32_t half, uint64_t total, uint32_t* start, uint32_t* vals) {
  const uint32_t M = half * (uint32_t)B, NP = M >> PART_LOG;
  void *pc, *rc;
  KZG_TRY(scratch_get(ctx, SCR_KEYS_A, ((size_t)NP * 3 + 16) * 4, &pc));
  KZG_TRY(scratch_get(ctx, SCR_KEYS_B, (total + 16) * 8, &rc));
  uint32_t* part_cnt = reinterpret_cast<uint32_t*>(pc);
  uint32_t* part_start = part_cnt + NP + 4;
  uint32_t* part_cur = part_start + NP + 4;
  KZG_CUDA_OK(ctx, cudaFuncSetAttribute(k_part_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(PART_STAGE * 4)));   // per device
  KZG_CUDA_OK(ctx, cudaMemsetAsync(part_cnt, 0, (size_t)NP * 4, ctx->stream));
  const dim3 grid(blocks_for(n, PART_TILE), (unsigned)B);
  k_part_count<<<grid, PART_TILE, 0, ctx->stream>>>(sp, n, c, W, half, part_cnt);
  KZG_LAUNCHED(ctx);
  k_part_scan<<<1, 1024, 0, ctx->stream>>>(part_cnt, NP, part_start, part_cur);
  KZG_LAUNCHED(ctx);
  k_part_scatter<<<grid, PART_TILE, 0, ctx->stream>>>(sp, n, c, W, n_stride, half, part_cur, reinterpret_cast<uint2*>(rc));
  KZG_LAUNCHED(ctx);
  k_part_sort<<<NP, 1024, PART_STAGE * 4, ctx->stream>>>(reinterpret_cast<const uint2*>(rc), part_start, NP, start, vals);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

// ---- 4. accumulate -----------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_l2(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// T = 2^logT threads share a bucket (contiguous slices of its sorted entries) so that the grid has
// ~2^18 threads with near-identical trip counts whatever the window size; partials[b*T + k] is
// folded by k_sum_chunks.  The next table point is prefetched into L2 while the current mixed
// addition (~3 500 instructions) runs.
// Buckets in order of decreasing length (one block: histogram of the lengths in shared memory, scan, scatter).  A thread
// of k_bucket_acc walks its slice of a bucket one addition after the other, so a warp takes as long as its LONGEST lane:
// with bucket lengths spread like a Poisson variable (34 +- 6 dense points per bucket at n = 2^17) the longest of 32
// lanes runs ~35 % more additions than the mean.  Handing out the buckets sorted by length gives every warp lanes of
// (nearly) equal length, and the long ones start first.
constexpr int ORDER_BINS = 2048;
__global__ void __launch_bounds__(1024)
k_bucket_order(const uint32_t* __restrict__ start, uint32_t M, uint32_t* __restrict__ perm) {
  __shared__ uint32_t hist[ORDER_BINS];
  __shared__ uint32_t warp_tot[32];
  for (uint32_t i = threadIdx.x; i < ORDER_BINS; i += 1024) hist[i] = 0;
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < M; b += 1024) {
    const uint32_t len = __ldg(start + b + 1) - __ldg(start + b);
    atomicAdd(&hist[ORDER_BINS - 1 - min(len, (uint32_t)ORDER_BINS - 1)], 1u);      // bin 0 = the longest buckets
  }
  __syncthreads();
  // exclusive scan of the 2048 bins: two per thread
  const uint32_t a0 = hist[2 * threadIdx.x], a1 = hist[2 * threadIdx.x + 1];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  uint32_t inc = a0 + a1;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += t; }
  if (lane == 31) warp_tot[w] = inc;
  __syncthreads();
  if (w == 0) {
    uint32_t x = warp_tot[lane], xi = x;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(0xffffffffu, xi, d); if (lane >= (uint32_t)d) xi += t; }
    warp_tot[lane] = xi - x;
  }
  __syncthreads();
  const uint32_t ex = warp_tot[w] + inc - (a0 + a1);
  hist[2 * threadIdx.x] = ex; hist[2 * threadIdx.x + 1] = ex + a0;
  __syncthreads();
  for (uint32_t b = threadIdx.x; b < M; b += 1024) {
    const uint32_t len = __ldg(start + b + 1) - __ldg(start + b);
    perm[atomicAdd(&hist[ORDER_BINS - 1 - min(len, (uint32_t)ORDER_BINS - 1)], 1u)] = b;
  }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB)
k_bucket_acc(const uint32_t* __restrict__ start, const uint32_t* __restrict__ vals,
             const char* __restrict__ tab, uint32_t pitch, uint32_t nbuckets, uint32_t logT, uint32_t heavy_cap,
             const uint32_t* __restrict__ perm,
             G1Xyzz* __restrict__ partials, uint32_t* __restrict__ heavy_list, uint32_t* __restrict__ heavy_count) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t b = t >> logT, k = t & ((1u << logT) - 1u);
  if (b >= nbuckets) return;
  if (perm) { b = __ldg(perm + b); t = (b << logT) | k; }      // slot of this thread's partial: by bucket, not by thread
  uint32_t s = start[b], e = start[b + 1];
  uint32_t len = e - s;
  G1Xyzz acc; set_inf(acc);
  if ((len >> logT) > heavy_cap) {
    if (k == 0) { uint32_t slot = atomicAdd(heavy_count, 1u); heavy_list[slot] = b; }
    partials[t] = acc;            // the heavy kernel overwrites slice 0 with the whole bucket
    return;
  }
  uint32_t lo = s + (uint32_t)(((uint64_t)len * k) >> logT);
  uint32_t hi = s + (uint32_t)(((uint64_t)len * (k + 1)) >> logT);
  uint32_t v = 0;
  if (lo < hi) {
    v = __ldg(vals + lo);
    const char* np = tab + (size_t)(v & 0x7FFFFFFFu) * pitch;
    prefetch_l2(np); prefetch_l2(np + 95);
  }
#pragma unroll 1
  for (uint32_t i = lo; i < hi; i++) {
    uint32_t vc = v;
    if (i + 1 < hi) {
      v = __ldg(vals + i + 1);
      const char* np = tab + (size_t)(v & 0x7FFFFFFFu) * pitch;
      prefetch_l2(np); prefetch_l2(np + 95);
    }
    G1Affine p; ld_affine_g(p, reinterpret_cast<const G1Affine*>(tab + (size_t)(vc & 0x7FFFFFFFu) * pitch));
    madd(acc, p, (vc >> 31) != 0);
  }
  partials[t] = acc;
}

// heavy buckets (adversarial inputs: all-equal scalars put n entries in one bucket): HEAVY_BLOCKS
// blocks share one bucket, each folds its strided entries through a shared-memory tree into
// heavy_parts[h*HEAVY_BLOCKS + blockIdx.x]; k_heavy_fold sums those into slice 0 of the bucket.
constexpr int HEAVY_BLOCKS = 32;

__global__ void __launch_bounds__(128)
k_bucket_heavy(const uint32_t* __restrict__ start, const uint32_t* __restrict__ vals,
               const char* __restrict__ tab, uint32_t pitch, G1Xyzz* __restrict__ heavy_parts,
               const uint32_t* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count) {
  __shared__ G1Xyzz sh[128];
  uint32_t nh = *heavy_count;
  for (uint32_t h = blockIdx.y; h < nh; h += gridDim.y) {
    uint32_t b = heavy_list[h];
    uint32_t s = start[b], e = start[b + 1];
    G1Xyzz acc; set_inf(acc);
#pragma unroll 1
    for (uint32_t k = s + blockIdx.x * blockDim.x + threadIdx.x; k < e; k += blockDim.x * HEAVY_BLOCKS) {
      uint32_t v = __ldg(vals + k);
      G1Affine p; ld_affine_g(p, reinterpret_cast<const G1Affine*>(tab + (size_t)(v & 0x7FFFFFFFu) * pitch));
      madd(acc, p, (v >> 31) != 0);
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
#pragma unroll 1
    for (uint32_t st = 64; st >= 1; st >>= 1) {
      if (threadIdx.x < st) { G1Xyzz o = sh[threadIdx.x + st]; add(acc, o); sh[threadIdx.x] = acc; }
      __syncthreads();
    }
    if (threadIdx.x == 0) heavy_parts[(size_t)h * HEAVY_BLOCKS + blockIdx.x] = acc;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(128)
k_heavy_fold(const G1Xyzz* __restrict__ heavy_parts, uint32_t logT, G1Xyzz* __restrict__ partials,
             const uint32_t* __restrict__ heavy_list, const uint32_t* __restrict__ heavy_count) {
  uint32_t nh = *heavy_count;
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < nh; h += gridDim.x * blockDim.x) {
    G1Xyzz acc; set_inf(acc);
    for (int k = 0; k < HEAVY_BLOCKS; k++) { G1Xyzz o = heavy_parts[(size_t)h * HEAVY_BLOCKS + k]; add(acc, o); }
    partials[(size_t)heavy_list[h] << logT] = acc;
  }
}

__global__ void k_set_inf(G1Xyzz* out, uint32_t count) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) { G1Xyzz p; set_inf(p); out[i] = p; }
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int32_t set_infinity(kzg_ctx* ctx, G1Xyzz* out_dev, uint32_t count) {
  k_set_inf<<<blocks_for(count ? count : 1, 128), 128, 0, ctx->stream>>>(out_dev, count);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

// count[0..M) -> start[0..M], cursor[0..M); the tile totals live behind cursor (the arena is sized for them)
static int32_t launch_bucket_scan(kzg_ctx* ctx, const uint32_t* count, uint32_t M, uint32_t* start, uint32_t* cursor) {
  const uint32_t ntiles = (M + SCAN_TILE - 1) / SCAN_TILE;
  if (ntiles > SCAN_TILE) return fail(ctx, KZG_E_BAD_ARG, "too many buckets for the counting sort");
  uint32_t* tile_tot = cursor + M + 2;
  k_bucket_tile_totals<<<ntiles, SCAN_TILE, 0, ctx->stream>>>(count, M, tile_tot);
  KZG_LAUNCHED(ctx);
  k_bucket_scan<<<ntiles, SCAN_TILE, 0, ctx->stream>>>(count, M, tile_tot, start, cursor);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

// affine halving rounds by problem size (measured, profiles/sweep_tree_rounds_r01.txt, sweep_lanes_r01.txt): one per
// doubling of the entry count from 2^20 entries on - 2 at n = 2^17, 5 at n = 2^20 (c = 16), at most 6.  A round's
// inversion chain is ~0.15 ms of latency; with several calls in flight (lanes) that latency is covered by the bulk
// kernels of the other calls, so one round more than the single-call break-even pays (n = 2^20, 4 lanes: 5.60 -> 5.51 ms;
// n = 2^17: 1.18 -> 1.10 ms) and costs a lone call < 1.5 %.
int auto_tree_rounds(uint64_t entries) {
  int r = 0;
  while (r < 6 && (entries >> (20 + r)) != 0) r++;
  return r;
}

// ---- level 2: the bucket sums reduced as a bucket problem of their own ----------------------------------------------
// sum_d d * B_d over 2^(c-1) buckets costs the running-sum tail ~80 Fq products per bucket (msm_tail.cu), which is what
// keeps the window at 17 bits.  Split the digit d = hi * 2^s + lo: sum_d d B_d = 2^s sum_hi hi H_hi + sum_lo lo L_lo with
// H_hi = sum of the B_d with that hi, L_lo likewise — two bucket problems of 2^(c-1-s) and 2^s - 1 buckets whose POINTS
// are the dense affine points the halving rounds of level 1 left (one or two per non-empty bucket), i.e. two affine
// additions of ~6 products per level-1 point instead of a share of the tail.  The points are addressed by position
// (pitch 96) exactly as level 1 addresses the table, so the same kernels run the level-2 halving rounds, the (now tiny)
// accumulate and the fused tail over 2 x calls sets of 2^(c-1-s) buckets; combine_level2 adds the two halves.
__global__ void __launch_bounds__(256)
k_l2_count(const uint32_t* __restrict__ seg, uint32_t Mtot, uint32_t half, uint32_t s_lo, uint32_t MS,
           uint32_t* __restrict__ cnt, uint32_t* __restrict__ off_lo) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Mtot) return;
  const uint32_t call = b / half, d = b % half + 1;
  const uint32_t hi = d >> s_lo, lo = d & ((1u << s_lo) - 1u);
  if (lo == 0) {
    // d = hi * 2^s: the first bucket of its hi group, whose points are the contiguous range up to d + 2^s (no atomics:
    // neighbouring buckets share hi and would serialise on one counter)
    const uint32_t end = min(b + (1u << s_lo), (call + 1) * half);
    cnt[(size_t)(call * 2) * MS + (hi - 1)] = __ldg(seg + end) - __ldg(seg + b);
    return;
  }
  const uint32_t len = __ldg(seg + b + 1) - __ldg(seg + b);
  if (len) off_lo[b] = atomicAdd(cnt + (size_t)(call * 2 + 1) * MS + (lo - 1), len);
}
__global__ void __launch_bounds__(256)
k_l2_fill(const uint32_t* __restrict__ seg, uint32_t Mtot, uint32_t half, uint32_t s_lo, uint32_t MS,
          const uint32_t* __restrict__ start2, const uint32_t* __restrict__ off_lo, uint32_t* __restrict__ vals2) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= Mtot) return;
  const uint32_t p0 = __ldg(seg + b), len = __ldg(seg + b + 1) - p0;
  if (!len) return;
  const uint32_t call = b / half, d = b % half + 1;
  const uint32_t hi = d >> s_lo, lo = d & ((1u << s_lo) - 1u);
  if (hi) {                                         // position inside the hi group = distance from the group's first point
    const uint32_t at = __ldg(start2 + (size_t)(call * 2) * MS + (hi - 1)) + (p0 - __ldg(seg + (b - lo)));
    for (uint32_t i = 0; i < len; i++) vals2[at + i] = p0 + i;
  }
  if (lo) {
    const uint32_t at = __ldg(start2 + (size_t)(call * 2 + 1) * MS + (lo - 1)) + off_lo[b];
    for (uint32_t i = 0; i < len; i++) vals2[at + i] = p0 + i;
  }
}

// seg[0..Mtot] = bounds of the dense points `dense` (at most E_bound of them) over Mtot = calls * half level-1 buckets
static int32_t finish_level2(kzg_ctx* ctx, const uint32_t* seg, const G1Affine* dense, uint64_t E_bound, uint32_t half,
                             int calls, int c, G1Xyzz* results_dev) {
  const uint32_t Mtot = half * (uint32_t)calls;
  const uint32_t s_lo = (uint32_t)(c - 1) / 2, MS = 1u << ((uint32_t)(c - 1) - s_lo);
  const uint32_t sets = 2u * (uint32_t)calls, M2 = sets * MS;
  const uint64_t E2 = 2 * E_bound;
  if (E2 >= (1ull << 31)) return fail(ctx, KZG_E_BAD_ARG, "second-level bucket problem too large");
  void *st2, *misc, *v2;
  KZG_TRY(scratch_get(ctx, SCR_L2_START, ((size_t)M2 + 3) * 4, &st2));
  KZG_TRY(scratch_get(ctx, SCR_L2_MISC, ((size_t)M2 * 2 + 8 + SCAN_TILE + (size_t)Mtot) * 4, &misc));
  KZG_TRY(scratch_get(ctx, SCR_L2_VALS, (E2 + 16) * 4, &v2));
  uint32_t* start2 = reinterpret_cast<uint32_t*>(st2);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(misc);
  uint32_t* cursor = cnt + M2 + 4;
  uint32_t* off = cursor + M2 + 2 + SCAN_TILE;
  uint32_t* vals2 = reinterpret_cast<uint32_t*>(v2);
  KZG_CUDA_OK(ctx, cudaMemsetAsync(cnt, 0, (size_t)M2 * 4, ctx->stream));
  k_l2_count<<<blocks_for(Mtot, 256), 256, 0, ctx->stream>>>(seg, Mtot, half, s_lo, MS, cnt, off);
  KZG_LAUNCHED(ctx);
  KZG_TRY(launch_bucket_scan(ctx, cnt, M2, start2, cursor));
  k_l2_fill<<<blocks_for(Mtot, 256), 256, 0, ctx->stream>>>(seg, Mtot, half, s_lo, MS, start2, off, vals2);
  KZG_LAUNCHED(ctx);

  // level-2 halving rounds until a bucket holds ~16 points (the accumulate then splits it over a few threads)
  int rounds2 = ctx->opt_l2_rounds;
  if (rounds2 < 0) {      // down to ~16 points per bucket, but no rounds of under 2^16 pairs (their inversion chain costs more than the accumulate)
    const uint64_t stop = std::max<uint64_t>(16ull * M2, 1ull << 17);
    rounds2 = 0;
    while (rounds2 < 14 && (E2 >> rounds2) > stop) rounds2++;
  }
  const uint32_t* acc_start = start2;
  const uint32_t* acc_vals = vals2;
  const char* acc_tab = reinterpret_cast<const char*>(dense);
  uint64_t acc_entries = E2;
  if (rounds2 > 0) {
    const G1Affine* d2 = nullptr;
    KZG_TRY(msm_tree_rounds(ctx, start2, vals2, reinterpret_cast<const char*>(dense), M2, E2, rounds2, &d2, &acc_start,
                            &acc_vals, &acc_entries, DENSE_PITCH, 2));
    acc_tab = reinterpret_cast<const char*>(d2);
  }
  uint32_t logT = 0;
  while (((uint64_t)M2 << logT) < (1u << 17) && acc_entries / ((uint64_t)M2 << logT) > 8 && logT < 6) logT++;
  void *parts, *hv, *hp, *r2;
  KZG_TRY(scratch_get(ctx, SCR_PARTIALS, ((size_t)M2 << logT) * sizeof(G1Xyzz), &parts));
  const uint32_t heavy_cap = (uint32_t)ctx->opt_heavy_cap;
  const size_t heavy_max = E2 / heavy_cap + 2 + 64;
  KZG_TRY(scratch_get(ctx, SCR_HEAVY, (heavy_max + 1) * 4, &hv));
  uint32_t* heavy_count = reinterpret_cast<uint32_t*>(hv);
  uint32_t* heavy_list = heavy_count + 1;
  KZG_CUDA_OK(ctx, cudaMemsetAsync(heavy_count, 0, 4, ctx->stream));
  const uint32_t* perm = nullptr;
  if (ctx->opt_acc_order) {
    void* pm;
    KZG_TRY(scratch_get(ctx, SCR_ORDER, (size_t)M2 * 4 + 16, &pm));
    k_bucket_order<<<1, 1024, 0, ctx->stream>>>(acc_start, M2, reinterpret_cast<uint32_t*>(pm));
    KZG_LAUNCHED(ctx);
    perm = reinterpret_cast<const uint32_t*>(pm);
  }
  k_bucket_acc<4><<<blocks_for((uint64_t)M2 << logT, 128), 128, 0, ctx->stream>>>(
      acc_start, acc_vals, acc_tab, DENSE_PITCH, M2, logT, heavy_cap, perm, reinterpret_cast<G1Xyzz*>(parts), heavy_list, heavy_count);
  KZG_LAUNCHED(ctx);
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  KZG_TRY(scratch_get(ctx, SCR_HEAVY_PARTS, heavy_max * HEAVY_BLOCKS * sizeof(G1Xyzz), &hp));
  k_bucket_heavy<<<dim3(HEAVY_BLOCKS, 16), 128, 0, ctx->stream>>>(acc_start, acc_vals, acc_tab, DENSE_PITCH,
                                                                 reinterpret_cast<G1Xyzz*>(hp), heavy_list, heavy_count);
  KZG_LAUNCHED(ctx);
  k_heavy_fold<<<8, 128, 0, ctx->stream>>>(reinterpret_cast<G1Xyzz*>(hp), logT, reinterpret_cast<G1Xyzz*>(parts), heavy_list, heavy_count);
  KZG_LAUNCHED(ctx);
  KZG_TRY(scratch_get(ctx, SCR_RED_A, ((size_t)sets + 4) * sizeof(G1Xyzz), &r2));
  KZG_TRY(reduce_buckets_multi(ctx, reinterpret_cast<G1Xyzz*>(parts), MS, logT, (int)sets, reinterpret_cast<G1Xyzz*>(r2)));
  KZG_TRY(combine_level2(ctx, reinterpret_cast<G1Xyzz*>(r2), (int)s_lo, calls, results_dev));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  return KZG_OK;
}

static inline bool use_level2(const kzg_ctx* ctx, int c) {
  return ctx->opt_msm_l2 == 1 || (ctx->opt_msm_l2 < 0 && c >= 18);
}
// halving rounds of level 1 when level 2 follows: until a bucket holds about one point
static inline int level1_rounds_for_l2(uint64_t total, uint32_t Mtot, int at_least) {
  int r = at_least > 1 ? at_least : 1;
  while (r < 12 && (total >> r) > (uint64_t)Mtot) r++;
  return r;
}

int32_t msm_bucket(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t n,
                   G1Xyzz* result_dev) {
  if (n == 0) return set_infinity(ctx, result_dev, 1);
  if (!s->msm_tab) return fail(ctx, KZG_E_BAD_ARG, "msm table not prepared");
  const int c = s->msm_c, W = s->msm_w;
  const uint32_t half = 1u << (c - 1);
  const uint64_t total = (uint64_t)W * n;
  if ((uint64_t)W * s->n >= (1ull << 31) || total >= (1ull << 32))
    return fail(ctx, KZG_E_BAD_ARG, "msm too large for 31-bit table indices");

  void *va, *vb, *st, *bk, *hv;
  KZG_TRY(scratch_get(ctx, SCR_VALS_B, total * 4, &vb));
  KZG_TRY(scratch_get(ctx, SCR_BUCKET_START, ((size_t)half + 3) * 4, &st));
  KZG_TRY(scratch_get(ctx, SCR_BUCKETS, (size_t)half * sizeof(G1Xyzz), &bk));
  const uint32_t heavy_cap = (uint32_t)ctx->opt_heavy_cap;
  size_t heavy_max = total / heavy_cap + 2 + 64;
  KZG_TRY(scratch_get(ctx, SCR_HEAVY, (heavy_max + 1) * 4, &hv));
  uint32_t* heavy_count = reinterpret_cast<uint32_t*>(hv);
  uint32_t* heavy_list = heavy_count + 1;

  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  // Counting sort at every size since the scan is tiled (two launches): 21 ps per entry at 31 M entries (two fused calls at
  // n = 2^20, profiles/launches_r02_fused_n20.csv) against 22 ps for recode + two radix passes + bounds at 16-bit keys —
  // and 17-bit keys (n >= 2^20) would cost the radix sort a third pass.  The cub path stays behind option msm_sort = 1.
  const bool counting = ctx->opt_msm_sort <= 0 || ctx->opt_msm_sort == 2;
  if (ctx->opt_msm_sort == 2 && part_sort_fits(half, 1)) {
    // two-level partition sort (option msm_sort = 2): shared-memory atomics instead of two global ones per entry
    ScalarPtrs sp = {}; sp.p[0] = scalars_dev;
    KZG_TRY(launch_part_sort(ctx, sp, 1, n, c, W, s->n, half, total, reinterpret_cast<uint32_t*>(st), reinterpret_cast<uint32_t*>(vb)));
  } else if (counting) {
    // counting sort: count -> scan -> scatter (see above); count[] and cursor[] share one small arena
    void* cc;
    KZG_TRY(scratch_get(ctx, SCR_KEYS_A, ((size_t)half * 2 + 8 + SCAN_TILE) * 4, &cc));
    uint32_t* count = reinterpret_cast<uint32_t*>(cc);
    uint32_t* cursor = count + half + 4;
    KZG_CUDA_OK(ctx, cudaMemsetAsync(count, 0, (size_t)half * 4, ctx->stream));
    ScalarPtrs sp = {}; sp.p[0] = scalars_dev;
    k_digit_count<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(sp, n, c, W, half, count);
    KZG_LAUNCHED(ctx);
    KZG_TRY(launch_bucket_scan(ctx, count, half, reinterpret_cast<uint32_t*>(st), cursor));
    k_digit_scatter<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(sp, n, c, W, s->n, half, cursor, reinterpret_cast<uint32_t*>(vb));
    KZG_LAUNCHED(ctx);
  } else {
    // radix path (option msm_sort = 1, kept for comparison): recode -> cub::DeviceRadixSort over c key bits -> bounds
    void *ka, *kb;
    KZG_TRY(scratch_get(ctx, SCR_KEYS_A, total * 4, &ka));
    KZG_TRY(scratch_get(ctx, SCR_KEYS_B, total * 4, &kb));
    KZG_TRY(scratch_get(ctx, SCR_VALS_A, total * 4, &va));
    k_recode<<<blocks_for(n, 256), 256, 0, ctx->stream>>>(scalars_dev, n, c, W, s->n,
                                                          reinterpret_cast<uint32_t*>(ka), reinterpret_cast<uint32_t*>(va));
    KZG_LAUNCHED(ctx);
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, reinterpret_cast<uint32_t*>(ka), reinterpret_cast<uint32_t*>(kb),
                                    reinterpret_cast<uint32_t*>(va), reinterpret_cast<uint32_t*>(vb), (uint64_t)total, 0, c,
                                    ctx->stream);
    void* tmp;
    KZG_TRY(scratch_get(ctx, SCR_SORT_TMP, tmp_bytes + 16, &tmp));
    KZG_CUDA_OK(ctx, cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, reinterpret_cast<uint32_t*>(ka),
                                                     reinterpret_cast<uint32_t*>(kb), reinterpret_cast<uint32_t*>(va),
                                                     reinterpret_cast<uint32_t*>(vb), (uint64_t)total, 0, c, ctx->stream));
    ctx->launches += 4;  // cub onesweep: histogram + scan + one pass per 8 key bits (approximate)
    k_bucket_bounds<<<blocks_for(total, 256), 256, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(kb), total, half,
                                                                     reinterpret_cast<uint32_t*>(st));
    KZG_LAUNCHED(ctx);
  }
  KZG_CUDA_OK(ctx, cudaMemsetAsync(heavy_count, 0, 4, ctx->stream));

  // optional affine halving rounds (msm_tree.cu): the accumulate kernel then walks dense affine points
  const uint32_t* acc_start = reinterpret_cast<uint32_t*>(st);
  const uint32_t* acc_vals = reinterpret_cast<uint32_t*>(vb);
  const char* acc_tab = s->msm_tab;            // table entries on the 128-byte pitch; dense 96-byte points after a round
  uint32_t acc_pitch = MSM_TAB_PITCH;
  uint64_t acc_entries = total;
  int rounds = ctx->opt_msm_tree_rounds >= 0 ? ctx->opt_msm_tree_rounds : auto_tree_rounds(total);
  const bool l2 = use_level2(ctx, c);
  if (l2) rounds = ctx->opt_msm_tree_rounds > 0 ? ctx->opt_msm_tree_rounds : level1_rounds_for_l2(total, half, rounds);
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  if (rounds > 0) {
    const G1Affine* dense = nullptr;
    KZG_TRY(msm_tree_rounds(ctx, acc_start, acc_vals, acc_tab, half, total, rounds, &dense, &acc_start, &acc_vals,
                            &acc_entries));
    acc_tab = reinterpret_cast<const char*>(dense); acc_pitch = DENSE_PITCH;
  }
  if (l2) return finish_level2(ctx, acc_start, reinterpret_cast<const G1Affine*>(acc_tab), acc_entries, half, 1, c, result_dev);

  // threads per bucket: aim for ~2^18 accumulate threads
  uint32_t logT = 0;
  if (ctx->opt_acc_split >= 0) logT = (uint32_t)ctx->opt_acc_split;
  else while ((half << logT) < (1u << 18) && (acc_entries >> (c - 1 + logT)) > (uint64_t)(rounds > 0 ? 8 : ctx->opt_acc_min_entries)) logT++;
  const uint32_t T = 1u << logT;
  void* parts = bk;
  if (logT) KZG_TRY(scratch_get(ctx, SCR_PARTIALS, ((size_t)half << logT) * sizeof(G1Xyzz), &parts));
  {
    // After the halving rounds this kernel folds only ~16 dense points per bucket: 2^16 threads = 512 blocks, which is
    // 1.15 waves at 3 blocks per SM but ONE wave at 4 (128 registers; n = 2^17 alone: 1.59 -> 1.55 ms).  Without rounds
    // (all W n additions here) 3 blocks per SM and 168 registers are faster (profiles/sweep_acc_r01.txt).
    auto kern = k_bucket_acc<3>;
    if (ctx->opt_acc_minblocks == 4 || (ctx->opt_acc_minblocks == 0 && rounds > 0)) kern = k_bucket_acc<4>;
    else if (ctx->opt_acc_minblocks == 5) kern = k_bucket_acc<5>;
    else if (ctx->opt_acc_minblocks == 6) kern = k_bucket_acc<6>;
    const uint32_t* perm = nullptr;
    if (ctx->opt_acc_order) {
      void* pm;
      KZG_TRY(scratch_get(ctx, SCR_ORDER, (size_t)half * 4 + 16, &pm));
      k_bucket_order<<<1, 1024, 0, ctx->stream>>>(acc_start, half, reinterpret_cast<uint32_t*>(pm));
      KZG_LAUNCHED(ctx);
      perm = reinterpret_cast<const uint32_t*>(pm);
    }
    if (rounds == 0) KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[0], ctx->stream));
    kern<<<blocks_for((uint64_t)half << logT, 128), 128, 0, ctx->stream>>>(
        acc_start, acc_vals, acc_tab, acc_pitch, half, logT, heavy_cap, perm,
        reinterpret_cast<G1Xyzz*>(parts), heavy_list, heavy_count);
  }
  KZG_LAUNCHED(ctx);
  if (rounds == 0) KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[1], ctx->stream));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  {
    void* hp;
    KZG_TRY(scratch_get(ctx, SCR_HEAVY_PARTS, heavy_max * HEAVY_BLOCKS * sizeof(G1Xyzz), &hp));
    k_bucket_heavy<<<dim3(HEAVY_BLOCKS, 16), 128, 0, ctx->stream>>>(
        acc_start, acc_vals, acc_tab, acc_pitch, reinterpret_cast<G1Xyzz*>(hp), heavy_list, heavy_count);
    KZG_LAUNCHED(ctx);
    k_heavy_fold<<<8, 128, 0, ctx->stream>>>(reinterpret_cast<G1Xyzz*>(hp), logT, reinterpret_cast<G1Xyzz*>(parts),
                                             heavy_list, heavy_count);
    KZG_LAUNCHED(ctx);
  }
  KZG_TRY(reduce_buckets(ctx, reinterpret_cast<G1Xyzz*>(parts), half, logT, result_dev));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  return KZG_OK;
}

// ---- several calls as ONE bucket problem -------------------------------------------------------------------------
// B sums over the same bases (calls that are in flight together, kzg_msm_g1_submit* with option msm_coalesce; the rows of
// a batch of long polynomials) are a single MSM with B times the buckets: entry (call, window j, point i) goes to bucket
// call * 2^(c-1) + |digit| - 1 and reads the SAME table entry j * n + i.  Sort, halving rounds and accumulation then run
// once over B * W * n entries — kernels B times the size, one inversion chain per round instead of B — and only the
// bucket reduction runs per call (one launch set, blockIdx.y = call).  At n = 2^17 (the per-GPU share of the 8-GPU MSM)
// the kernels of a lone call are single partial waves; four calls fused run with the efficiency of n = 2^19.
int32_t msm_bucket_multi(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* const* scalars_dev, int B, uint64_t n,
                         G1Xyzz* results_dev) {
  if (B < 1 || B > KZG_MAX_COALESCE) return fail(ctx, KZG_E_BAD_ARG, "bad number of fused calls");
  if (B == 1) return msm_bucket(ctx, s, scalars_dev[0], n, results_dev);
  if (n == 0) return set_infinity(ctx, results_dev, (uint32_t)B);
  if (!s->msm_tab) return fail(ctx, KZG_E_BAD_ARG, "msm table not prepared");
  const int c = s->msm_c, W = s->msm_w;
  const uint32_t half = 1u << (c - 1);
  const uint32_t M = half * (uint32_t)B;
  const uint64_t total = (uint64_t)W * n * (uint64_t)B;
  if ((uint64_t)W * s->n >= (1ull << 31) || total >= (1ull << 32))
    return fail(ctx, KZG_E_BAD_ARG, "msm too large for 31-bit table indices");
  if ((uint64_t)M > (1ull << 20)) return fail(ctx, KZG_E_BAD_ARG, "fused calls: too many buckets (window too wide)");

  void *vb, *st, *hv, *cc;
  KZG_TRY(scratch_get(ctx, SCR_VALS_B, total * 4, &vb));
  KZG_TRY(scratch_get(ctx, SCR_BUCKET_START, ((size_t)M + 3) * 4, &st));
  const uint32_t heavy_cap = (uint32_t)ctx->opt_heavy_cap;
  size_t heavy_max = total / heavy_cap + 2 + 64;
  KZG_TRY(scratch_get(ctx, SCR_HEAVY, (heavy_max + 1) * 4, &hv));
  uint32_t* heavy_count = reinterpret_cast<uint32_t*>(hv);
  uint32_t* heavy_list = heavy_count + 1;
  KZG_TRY(scratch_get(ctx, SCR_KEYS_A, ((size_t)M * 2 + 8 + SCAN_TILE) * 4, &cc));
  uint32_t* count = reinterpret_cast<uint32_t*>(cc);
  uint32_t* cursor = count + M + 4;

  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[2], ctx->stream));
  ScalarPtrs sp = {};
  for (int b = 0; b < B; b++) sp.p[b] = scalars_dev[b];
  if (ctx->opt_msm_sort == 2 && part_sort_fits(half, B)) {
    KZG_TRY(launch_part_sort(ctx, sp, B, n, c, W, s->n, half, total, reinterpret_cast<uint32_t*>(st), reinterpret_cast<uint32_t*>(vb)));
  } else {
    KZG_CUDA_OK(ctx, cudaMemsetAsync(count, 0, (size_t)M * 4, ctx->stream));
    k_digit_count<<<dim3(blocks_for(n, 256), (unsigned)B), 256, 0, ctx->stream>>>(sp, n, c, W, half, count);
    KZG_LAUNCHED(ctx);
    KZG_TRY(launch_bucket_scan(ctx, count, M, reinterpret_cast<uint32_t*>(st), cursor));
    k_digit_scatter<<<dim3(blocks_for(n, 256), (unsigned)B), 256, 0, ctx->stream>>>(sp, n, c, W, s->n, half, cursor,
                                                                                   reinterpret_cast<uint32_t*>(vb));
    KZG_LAUNCHED(ctx);
  }
  KZG_CUDA_OK(ctx, cudaMemsetAsync(heavy_count, 0, 4, ctx->stream));

  const uint32_t* acc_start = reinterpret_cast<uint32_t*>(st);
  const uint32_t* acc_vals = reinterpret_cast<uint32_t*>(vb);
  const char* acc_tab = s->msm_tab;
  uint32_t acc_pitch = MSM_TAB_PITCH;
  uint64_t acc_entries = total;
  int rounds = ctx->opt_msm_tree_rounds >= 0 ? ctx->opt_msm_tree_rounds : auto_tree_rounds(total);
  const bool l2 = use_level2(ctx, c);
  if (l2) rounds = ctx->opt_msm_tree_rounds > 0 ? ctx->opt_msm_tree_rounds : level1_rounds_for_l2(total, M, rounds);
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  if (rounds > 0) {
    const G1Affine* dense = nullptr;
    KZG_TRY(msm_tree_rounds(ctx, acc_start, acc_vals, acc_tab, M, total, rounds, &dense, &acc_start, &acc_vals, &acc_entries));
    acc_tab = reinterpret_cast<const char*>(dense); acc_pitch = DENSE_PITCH;
  }
  if (l2) return finish_level2(ctx, acc_start, reinterpret_cast<const G1Affine*>(acc_tab), acc_entries, half, B, c, results_dev);
  uint32_t logT = 0;
  if (ctx->opt_acc_split >= 0) logT = (uint32_t)ctx->opt_acc_split;
  else while (((uint64_t)M << logT) < (1u << 18) && acc_entries / ((uint64_t)M << logT) > (uint64_t)(rounds > 0 ? 8 : ctx->opt_acc_min_entries)) logT++;
  void* parts;
  KZG_TRY(scratch_get(ctx, SCR_PARTIALS, ((size_t)M << logT) * sizeof(G1Xyzz), &parts));
  {
    auto kern = k_bucket_acc<3>;
    if (ctx->opt_acc_minblocks == 4 || (ctx->opt_acc_minblocks == 0 && rounds > 0)) kern = k_bucket_acc<4>;
    else if (ctx->opt_acc_minblocks == 5) kern = k_bucket_acc<5>;
    else if (ctx->opt_acc_minblocks == 6) kern = k_bucket_acc<6>;
    const uint32_t* perm = nullptr;
    if (ctx->opt_acc_order) {
      void* pm;
      KZG_TRY(scratch_get(ctx, SCR_ORDER, (size_t)M * 4 + 16, &pm));
      k_bucket_order<<<1, 1024, 0, ctx->stream>>>(acc_start, M, reinterpret_cast<uint32_t*>(pm));
      KZG_LAUNCHED(ctx);
      perm = reinterpret_cast<const uint32_t*>(pm);
    }
    if (rounds == 0) KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[0], ctx->stream));
    kern<<<blocks_for((uint64_t)M << logT, 128), 128, 0, ctx->stream>>>(
        acc_start, acc_vals, acc_tab, acc_pitch, M, logT, heavy_cap, perm,
        reinterpret_cast<G1Xyzz*>(parts), heavy_list, heavy_count);
    KZG_LAUNCHED(ctx);
    if (rounds == 0) KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev_k[1], ctx->stream));
  }
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  {
    void* hp;
    KZG_TRY(scratch_get(ctx, SCR_HEAVY_PARTS, heavy_max * HEAVY_BLOCKS * sizeof(G1Xyzz), &hp));
    k_bucket_heavy<<<dim3(HEAVY_BLOCKS, 16), 128, 0, ctx->stream>>>(
        acc_start, acc_vals, acc_tab, acc_pitch, reinterpret_cast<G1Xyzz*>(hp), heavy_list, heavy_count);
    KZG_LAUNCHED(ctx);
    k_heavy_fold<<<8, 128, 0, ctx->stream>>>(reinterpret_cast<G1Xyzz*>(hp), logT, reinterpret_cast<G1Xyzz*>(parts),
                                             heavy_list, heavy_count);
    KZG_LAUNCHED(ctx);
  }
  KZG_TRY(reduce_buckets_multi(ctx, reinterpret_cast<G1Xyzz*>(parts), half, logT, B, results_dev));
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[3], ctx->stream));
  return KZG_OK;
}

}  // namespace kzg
