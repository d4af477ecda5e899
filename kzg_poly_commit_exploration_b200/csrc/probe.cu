// Roofline probes and the on-device field self-test.
//   probe_imad   : independent 32x32->64 products (IMAD.WIDE.U32) on every SM -> measured issue rate of the multiplier pipe
//                  (the nominal 64 lanes/clk/SM must be replaced by a measurement)
//   probe_fq_mul : sustained throughput of the shipped Fq Montgomery multiplication
//   selftest     : canonical products a*b mod q computed through to_mont / mul / from_mont
#include "engine.cuh"

namespace kzg {

constexpr int PROBE_ITERS = 2048;
constexpr int PROBE_CHAINS = 8;

// The multiplier pipe on its own: 8 independent chains per thread of p <- lo(p) * hi(p) (32 x 32 -> 64): every step is
// exactly ONE IMAD.WIDE.U32 (zero addend) whose two result halves are the next step's operands, so there is nothing for
// ptxas to fold and no companion instruction on another pipe.  History: round 1's probe accumulated the same product four
// times (`mad.wide.u32 acc, a, b, acc`), ptxas folded three of the four away and split the accumulation into IADD3 pairs,
// so it timed one IMAD.WIDE plus eight adds per "four multiply-adds"; a first repair consumed each product with a LOP3,
// and measured IMAD.WIDE + LOP3 back to back (ALU instructions do not hide under wide multiplies on this chip,
// profiles/ubench_imad_r01.txt): 30.6 lanes/clk/SM, BELOW the shipped field product.  SASS of this loop body
// (cuobjdump, profiles/probe_imad_r02_sass.txt): 32 x IMAD.WIDE.U32 Rp, Rlo, Rhi, RZ + loop overhead.
// The additions of a real multiply-ACCUMULATE come on top (IADD3 pairs on the ALU pipe, or fused as the carry-chained
// IMAD.WIDE.U32.X the field code uses, which holds the pipe longer: kzg_probe_fq_mul measures that one).
__global__ void __launch_bounds__(256) k_probe_imad(uint32_t* out, uint32_t seed) {
  uint64_t p[PROBE_CHAINS];
#pragma unroll
  for (int k = 0; k < PROBE_CHAINS; k++)
    p[k] = ((uint64_t)(seed + threadIdx.x * 2654435761u + (uint32_t)k * 0x85EBCA6Bu) << 32) | ((seed ^ 0x9E3779B9u) + blockIdx.x * 977u + (uint32_t)k);
#pragma unroll 1
  for (int it = 0; it < PROBE_ITERS; it++) {
#pragma unroll
    for (int rep = 0; rep < 4; rep++) {
#pragma unroll
      for (int k = 0; k < PROBE_CHAINS; k++) {
        const uint32_t lo = (uint32_t)p[k], hi = (uint32_t)(p[k] >> 32);
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p[k]) : "r"(lo), "r"(hi));
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < PROBE_CHAINS; k++) s ^= p[k];
  if (s == 0x1234567ull) out[0] = (uint32_t)s;   // never true in practice; keeps the chains live
}

__global__ void __launch_bounds__(128) k_probe_fq_mul(Fq* io, int iters) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  Fq a = io[(2 * t) & 1023], b = io[(2 * t + 1) & 1023];
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    Fq c; mul(c, a, b);
    a = b; b = c;
  }
  if (is_zero(b) && is_zero(a)) io[0] = b;  // keep live; never true for nonzero inputs
}

__global__ void k_fill_fq(Fq* io, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x;
#pragma unroll
  for (int k = 0; k < 12; k++) x.v[k] = (i + 1) * 2654435761u + k * 0x9E3779B9u;
  x.v[11] &= 0x0FFFFFFFu;
  io[i] = x;
}

__global__ void k_selftest_fq_mul(const uint8_t* a_le, const uint8_t* b_le, uint64_t n, uint8_t* out_le) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq a, b;
  const uint32_t* aw = reinterpret_cast<const uint32_t*>(a_le + i * 48);
  const uint32_t* bw = reinterpret_cast<const uint32_t*>(b_le + i * 48);
#pragma unroll
  for (int k = 0; k < 12; k++) { a.v[k] = aw[k]; b.v[k] = bw[k]; }
  Fq am, bm, cm, c, s, d;
  to_mont(am, a); to_mont(bm, b);
  mul(cm, am, bm);
  sqr(s, am); mul(s, s, bm);     // a^2 b
  add(d, cm, s); sub(d, d, am);  // ab + a^2 b - a
  from_mont(c, d);
  uint32_t* ow = reinterpret_cast<uint32_t*>(out_le + i * 48);
#pragma unroll
  for (int k = 0; k < 12; k++) ow[k] = c.v[k];
}

static int32_t time_kernel(kzg_ctx* ctx, float* ms) {
  KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[1], ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  KZG_CUDA_OK(ctx, cudaEventElapsedTime(ms, ctx->ev[0], ctx->ev[1]));
  return KZG_OK;
}

int32_t probe_imad(kzg_ctx* ctx, double* wide_mac_per_s) {
  void* m;
  KZG_TRY(scratch_get(ctx, SCR_MISC, 1024, &m));
  const int blocks = ctx->sm_count * 8, threads = 256;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    k_probe_imad<<<blocks, threads, 0, ctx->stream>>>(reinterpret_cast<uint32_t*>(m), 12345u + rep);
    KZG_LAUNCHED(ctx);
    float ms = 0;
    KZG_TRY(time_kernel(ctx, &ms));
    double ops = (double)blocks * threads * PROBE_ITERS * 4.0 * PROBE_CHAINS;
    double rate = ops / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  *wide_mac_per_s = best;
  return KZG_OK;
}

int32_t probe_fq_mul(kzg_ctx* ctx, double* fq_mul_per_s) {
  void* m;
  KZG_TRY(scratch_get(ctx, SCR_XYZZ_TMP, 1024 * sizeof(Fq), &m));
  k_fill_fq<<<8, 128, 0, ctx->stream>>>(reinterpret_cast<Fq*>(m), 1024);
  KZG_LAUNCHED(ctx);
  const int blocks = ctx->sm_count * 16, threads = 128, iters = 2000;
  double best = 0;
  for (int rep = 0; rep < 4; rep++) {
    KZG_CUDA_OK(ctx, cudaEventRecord(ctx->ev[0], ctx->stream));
    k_probe_fq_mul<<<blocks, threads, 0, ctx->stream>>>(reinterpret_cast<Fq*>(m), iters);
    KZG_LAUNCHED(ctx);
    float ms = 0;
    KZG_TRY(time_kernel(ctx, &ms));
    double rate = (double)blocks * threads * iters / (ms * 1e-3);
    if (rep > 0 && rate > best) best = rate;
  }
  *fq_mul_per_s = best;
  return KZG_OK;
}

}  // namespace kzg

extern "C" int32_t kzg_selftest_fq_mul(kzg_ctx* ctx, const uint8_t* a_le48, const uint8_t* b_le48, uint64_t n,
                                       uint8_t* out_le48) {
  using namespace kzg;
  if (!ctx || !a_le48 || !b_le48 || !out_le48) return KZG_E_BAD_ARG;
  if (n == 0) return KZG_OK;
  int prev = -1; cudaGetDevice(&prev); cudaSetDevice(ctx->device);
  void *a, *b, *o;
  KZG_TRY(scratch_get(ctx, SCR_KEYS_A, n * 48, &a));
  KZG_TRY(scratch_get(ctx, SCR_KEYS_B, n * 48, &b));
  KZG_TRY(scratch_get(ctx, SCR_VALS_A, n * 48, &o));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(a, a_le48, n * 48, cudaMemcpyHostToDevice, ctx->stream));
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(b, b_le48, n * 48, cudaMemcpyHostToDevice, ctx->stream));
  k_selftest_fq_mul<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(
      reinterpret_cast<uint8_t*>(a), reinterpret_cast<uint8_t*>(b), n, reinterpret_cast<uint8_t*>(o));
  KZG_LAUNCHED(ctx);
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(out_le48, o, n * 48, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (prev >= 0) cudaSetDevice(prev);
  return KZG_OK;
}
