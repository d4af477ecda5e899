// K5 — fixed-base commits for batches of small polynomials (EIP-4844-sized blobs).
//
// A blob MSM (n <= 4096 points) is far too small to fill 148 SMs with a bucket method, and every
// blob uses the SAME bases, so the setup carries a lookup table
//     fix[((i*W + j) << (c-1)) + d-1] = d * 2^(c j) * P_i ,  d = 1 .. 2^(c-1)
// (built once per setup, setup.cu).  A commit is then W signed-digit lookups per coefficient and
// nothing else: no buckets, no sort, no window doublings.  One block per (blob, split): a thread
// walks coefficients i = t, t+128, ... and their W digits, XYZZ += affine (8M+2S); the block
// folds its 128 partial sums through shared memory.
// Replaces the per-blob Polynomial::commit loop (src/polynomial.rs:207-212) and the commit half
// of Evaluation::generate_proof (:268).
#include "engine.cuh"

namespace kzg {

constexpr int FB_THREADS = 128;

__device__ __forceinline__ void ld_affine_b(G1Affine& o, const G1Affine* p) {
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

__global__ void __launch_bounds__(FB_THREADS)
k_fixed_batch(const uint8_t* __restrict__ scalars, uint64_t n, uint64_t stride, const G1Affine* __restrict__ fix,
              int c, int W, uint32_t splits, G1Xyzz* __restrict__ out) {
  __shared__ G1Xyzz sh[FB_THREADS];
  const uint32_t blob = blockIdx.y, sp = blockIdx.x;
  const uint8_t* sb = scalars + (size_t)blob * stride * 32;
  const uint32_t half = 1u << (c - 1);
  G1Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (uint64_t i = (uint64_t)sp * FB_THREADS + threadIdx.x; i < n; i += (uint64_t)splits * FB_THREADS) {
    Fr k;
    {
      const uint4* q = reinterpret_cast<const uint4*>(sb + i * 32);
      uint4 a = __ldg(q), b = __ldg(q + 1);
      k.v[0] = a.x; k.v[1] = a.y; k.v[2] = a.z; k.v[3] = a.w;
      k.v[4] = b.x; k.v[5] = b.y; k.v[6] = b.z; k.v[7] = b.w;
    }
    fr_canon_reduce(k);
    SignedDigits sd; sd.init(k, c);
    const G1Affine* row = fix + (((size_t)i * W) << (c - 1));
#pragma unroll 1
    for (int j = 0; j < W; j++) {
      bool neg;
      const uint32_t d = sd.next(j, &neg);
      if (d != 0) {
        G1Affine p; ld_affine_b(p, row + ((size_t)j << (c - 1)) + (d - 1));
        madd(acc, p, neg);
      }
    }
  }
  sh[threadIdx.x] = acc;
  __syncthreads();
#pragma unroll 1
  for (uint32_t st = FB_THREADS / 2; st >= 1; st >>= 1) {
    if (threadIdx.x < st) { G1Xyzz o = sh[threadIdx.x + st]; add(acc, o); sh[threadIdx.x] = acc; }
    __syncthreads();
  }
  if (threadIdx.x == 0) out[(size_t)blob * splits + sp] = acc;
}

// out[b] = sum of in[b*L .. b*L + L)
__global__ void __launch_bounds__(128)
k_sum_groups(const G1Xyzz* __restrict__ in, uint64_t groups, uint32_t L, G1Xyzz* __restrict__ out) {
  uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= groups) return;
  G1Xyzz acc; set_inf(acc);
  for (uint32_t l = 0; l < L; l++) { G1Xyzz b = in[g * L + l]; add(acc, b); }
  out[g] = acc;
}

int32_t msm_fixed_batch(kzg_ctx* ctx, const kzg_setup* s, const uint8_t* scalars_dev, uint64_t blobs,
                        uint64_t n, uint64_t stride_scalars, G1Xyzz* out_dev) {
  if (blobs == 0) return KZG_OK;
  if (!s->fix_tab || n > s->fix_n) return fail(ctx, KZG_E_BAD_ARG, "fixed-base table not prepared for this size");
  if (n > 0 && s->fix_w <= 32 && blobs >= (uint64_t)ctx->opt_tree_min_rows)
    return msm_fixed_batch_tree(ctx, s, scalars_dev, blobs, n, stride_scalars, out_dev);
  if (blobs > 65535) return fail(ctx, KZG_E_BAD_ARG, "more than 65535 blobs per call");
  // enough blocks to fill the machine: small batches split each MSM across blocks
  uint32_t splits = 1;
  uint64_t want_blocks = (uint64_t)ctx->sm_count * 4;
  if (blobs < want_blocks) {
    uint64_t max_splits = (n + FB_THREADS - 1) / FB_THREADS;
    splits = (uint32_t)((want_blocks + blobs - 1) / blobs);
    if (splits > max_splits) splits = (uint32_t)(max_splits ? max_splits : 1);
  }
  G1Xyzz* dst = out_dev;
  if (splits > 1) {
    void* tmp;
    KZG_TRY(scratch_get(ctx, SCR_RED_A, blobs * splits * sizeof(G1Xyzz), &tmp));
    dst = reinterpret_cast<G1Xyzz*>(tmp);
  }
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
  dim3 grid(splits, (unsigned)blobs);
  k_fixed_batch<<<grid, FB_THREADS, 0, ctx->stream>>>(scalars_dev, n, stride_scalars, s->fix_tab, s->fix_c, s->fix_w,
                                                      splits, dst);
  KZG_LAUNCHED(ctx);
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  if (splits > 1) {
    k_sum_groups<<<(unsigned)((blobs + 127) / 128), 128, 0, ctx->stream>>>(dst, blobs, splits, out_dev);
    KZG_LAUNCHED(ctx);
  }
  return KZG_OK;
}

}  // namespace kzg
