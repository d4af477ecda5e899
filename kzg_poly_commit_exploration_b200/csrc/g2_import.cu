// Device-side validation of imported G2 points (kzg_setup_import_g2 with 64+ points).
//
// Own translation unit with the field product compiled OUT OF LINE (KZG_FQ_NOINLINE, like msm_tail.cu): the Fq2 square
// root is two 381-bit exponentiations, and with the product inlined ptxas spends ~5 minutes on this one kernel; as a
// function call it compiles in seconds and the kernel is nowhere near a hot path.
#define KZG_FQ_NOINLINE 1
#include "engine.cuh"
#include "g2.cuh"

namespace kzg {

// Deserialize for G2Point x n on the device (src/curves.rs:273-343, blst_p2_uncompress semantics: flags,
// coordinates < q, Fq2 square root, on-curve; no subgroup check): bad = index of the first malformed point
__global__ void __launch_bounds__(64)
k_g2_validate(const uint8_t* __restrict__ in_c96, uint64_t n, unsigned long long* __restrict__ bad) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  __align__(16) uint8_t buf[96];      // filled with 128-bit stores
  const uint4* q = reinterpret_cast<const uint4*>(in_c96 + i * 96);
#pragma unroll
  for (int k = 0; k < 6; k++) reinterpret_cast<uint4*>(buf)[k] = __ldg(q + k);
  G2Affine a;
  if (uncompress(a, buf)) atomicMin(bad, (unsigned long long)i);
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int32_t validate_g2_points(kzg_ctx* ctx, const uint8_t* in_c96_dev, uint64_t n, int64_t* bad_index) {
  *bad_index = -1;
  if (n == 0) return KZG_OK;
  void* m;
  KZG_TRY(scratch_get(ctx, SCR_MISC, 256, &m));
  unsigned long long init = ~0ull;
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(m, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
  k_g2_validate<<<blocks_for(n, 64), 64, 0, ctx->stream>>>(in_c96_dev, n, reinterpret_cast<unsigned long long*>(m));
  KZG_LAUNCHED(ctx);
  unsigned long long bad = 0;
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(&bad, m, 8, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (bad != ~0ull) *bad_index = (int64_t)bad;
  return KZG_OK;
}

}  // namespace kzg
