// Integer-pipe microbenchmarks for sm_100a (run under gpurun): what does an IMAD.WIDE really cost?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench tools/ubench/ubench.cu && gpurun_out/ubench
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../kzg_poly_commit_exploration_b200/csrc/g1.cuh"
using namespace kzg;

#define ITERS 1024

// T1: independent 64-bit mad.wide chains, NCH per thread
template <int NCH>
__global__ void __launch_bounds__(256) t_wide_indep(uint32_t* out, uint32_t seed) {
  uint64_t acc[NCH]; uint32_t a[NCH];
  uint32_t b = seed + blockIdx.x;
#pragma unroll
  for (int k = 0; k < NCH; k++) { acc[k] = threadIdx.x + k; a[k] = seed * (k + 3) + threadIdx.x; }
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++)
#pragma unroll
      for (int k = 0; k < NCH; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(b));
    b += 7;
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= acc[k];
  if (s == 0x123456789ull) out[0] = 1;
}

// T2: carry chains: NCH independent chains of mad.lo.cc/madc.hi.cc pairs over 12 limbs (one "row" each)
template <int NCH>
__global__ void __launch_bounds__(256) t_wide_carry(uint32_t* out, uint32_t seed) {
  uint32_t acc[NCH][12]; uint32_t a[6];
  uint32_t b = seed + blockIdx.x;
#pragma unroll
  for (int k = 0; k < NCH; k++)
#pragma unroll
    for (int j = 0; j < 12; j++) acc[k][j] = threadIdx.x + k * 13 + j;
#pragma unroll
  for (int j = 0; j < 6; j++) a[j] = seed * (j + 3) + threadIdx.x;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) {
      asm volatile(
          "mad.lo.cc.u32 %0, %12, %18, %0;\n\tmadc.hi.cc.u32 %1, %12, %18, %1;\n\t"
          "madc.lo.cc.u32 %2, %13, %18, %2;\n\tmadc.hi.cc.u32 %3, %13, %18, %3;\n\t"
          "madc.lo.cc.u32 %4, %14, %18, %4;\n\tmadc.hi.cc.u32 %5, %14, %18, %5;\n\t"
          "madc.lo.cc.u32 %6, %15, %18, %6;\n\tmadc.hi.cc.u32 %7, %15, %18, %7;\n\t"
          "madc.lo.cc.u32 %8, %16, %18, %8;\n\tmadc.hi.cc.u32 %9, %16, %18, %9;\n\t"
          "madc.lo.cc.u32 %10, %17, %18, %10;\n\tmadc.hi.u32 %11, %17, %18, %11;\n\t"
          : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3]), "+r"(acc[k][4]), "+r"(acc[k][5]),
            "+r"(acc[k][6]), "+r"(acc[k][7]), "+r"(acc[k][8]), "+r"(acc[k][9]), "+r"(acc[k][10]), "+r"(acc[k][11])
          : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(b + k));
    }
    b += 7;
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++)
#pragma unroll
    for (int j = 0; j < 12; j++) s ^= acc[k][j];
  if (s == 0x12345678u) out[0] = 1;
}

// T3: same carry chain but multiplier operand is an immediate (as in the m*p half of the reduction)
template <int NCH>
__global__ void __launch_bounds__(256) t_wide_carry_imm(uint32_t* out, uint32_t seed) {
  uint32_t acc[NCH][12];
  uint32_t b = seed + blockIdx.x;
#pragma unroll
  for (int k = 0; k < NCH; k++)
#pragma unroll
    for (int j = 0; j < 12; j++) acc[k][j] = threadIdx.x + k * 13 + j;
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) {
      asm volatile(
          "mad.lo.cc.u32 %0, 0xffffaaab, %12, %0;\n\tmadc.hi.cc.u32 %1, 0xffffaaab, %12, %1;\n\t"
          "madc.lo.cc.u32 %2, 0xb153ffff, %12, %2;\n\tmadc.hi.cc.u32 %3, 0xb153ffff, %12, %3;\n\t"
          "madc.lo.cc.u32 %4, 0xf6b0f624, %12, %4;\n\tmadc.hi.cc.u32 %5, 0xf6b0f624, %12, %5;\n\t"
          "madc.lo.cc.u32 %6, 0xf38512bf, %12, %6;\n\tmadc.hi.cc.u32 %7, 0xf38512bf, %12, %7;\n\t"
          "madc.lo.cc.u32 %8, 0x434bacd7, %12, %8;\n\tmadc.hi.cc.u32 %9, 0x434bacd7, %12, %9;\n\t"
          "madc.lo.cc.u32 %10, 0x397fe69a, %12, %10;\n\tmadc.hi.u32 %11, 0x397fe69a, %12, %11;\n\t"
          : "+r"(acc[k][0]), "+r"(acc[k][1]), "+r"(acc[k][2]), "+r"(acc[k][3]), "+r"(acc[k][4]), "+r"(acc[k][5]),
            "+r"(acc[k][6]), "+r"(acc[k][7]), "+r"(acc[k][8]), "+r"(acc[k][9]), "+r"(acc[k][10]), "+r"(acc[k][11])
          : "r"(b + k));
    }
    b += 7;
  }
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++)
#pragma unroll
    for (int j = 0; j < 12; j++) s ^= acc[k][j];
  if (s == 0x12345678u) out[0] = 1;
}

// T4: the shipped Fq multiplication, NCH independent products per thread per iteration
template <int NCH>
__global__ void __launch_bounds__(128) t_fq_mul(Fq* io, int iters) {
  Fq a[NCH], b[NCH];
#pragma unroll
  for (int k = 0; k < NCH; k++) { a[k] = io[(threadIdx.x + k) & 1023]; b[k] = io[(threadIdx.x * 3 + k + 1) & 1023]; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int k = 0; k < NCH; k++) { Fq c; mul(c, a[k], b[k]); a[k] = b[k]; b[k] = c; }
  }
  bool z = true;
#pragma unroll
  for (int k = 0; k < NCH; k++) z = z && is_zero(b[k]);
  if (z) io[0] = b[0];
}

// T5: XYZZ mixed additions back to back (the accumulate inner loop without memory)
__global__ void __launch_bounds__(128) t_madd(G1Affine* pts, G1Xyzz* out, int iters) {
  G1Affine p = pts[threadIdx.x & 255];
  G1Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (int it = 0; it < iters; it++) { madd(acc, p, (it & 1) != 0); p.x.v[0] ^= acc.x.v[3] & 1; }
  if (is_zero(acc.x) && is_zero(acc.y)) out[0] = acc;
}

__global__ void k_fill(Fq* io, uint32_t n) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fq x;
  for (int k = 0; k < 12; k++) x.v[k] = (i + 1) * 2654435761u + k * 0x9E3779B9u;
  x.v[11] &= 0x0FFFFFFFu;
  io[i] = x;
}

template <class F>
static double timeit(F launch, int reps = 3) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    cudaEventRecord(e0); launch(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;
  }
  return best * 1e-3;
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  double clk = 1.965e9;
  uint32_t* out; cudaMalloc(&out, 1024);
  Fq* io; cudaMalloc(&io, 4096 * sizeof(Fq)); k_fill<<<32, 128>>>(io, 4096);
  printf("SMs %d\n", sms);
  auto report = [&](const char* name, double ops, double s) {
    printf("%-34s %.3e ops/s  %.2f lanes/clk/SM  (%.2f cyc per warp-instr per SMSP)\n", name, ops / s, ops / s / sms / clk,
           32.0 * 4 * sms * clk / (ops / s));
  };
  for (int bps : {2, 4, 8}) {
    int blocks = sms * bps;
    printf("-- %d blocks/SM of 256 threads\n", bps);
    report("wide indep x4", (double)blocks * 256 * ITERS * 16, timeit([&] { t_wide_indep<4><<<blocks, 256>>>(out, 1); }));
    report("wide indep x8", (double)blocks * 256 * ITERS * 32, timeit([&] { t_wide_indep<8><<<blocks, 256>>>(out, 1); }));
    report("wide carry-chain x1", (double)blocks * 256 * ITERS * 6, timeit([&] { t_wide_carry<1><<<blocks, 256>>>(out, 1); }));
    report("wide carry-chain x2", (double)blocks * 256 * ITERS * 12, timeit([&] { t_wide_carry<2><<<blocks, 256>>>(out, 1); }));
    report("wide carry-chain x4", (double)blocks * 256 * ITERS * 24, timeit([&] { t_wide_carry<4><<<blocks, 256>>>(out, 1); }));
    report("wide carry-chain imm x2", (double)blocks * 256 * ITERS * 12, timeit([&] { t_wide_carry_imm<2><<<blocks, 256>>>(out, 1); }));
    report("wide carry-chain imm x4", (double)blocks * 256 * ITERS * 24, timeit([&] { t_wide_carry_imm<4><<<blocks, 256>>>(out, 1); }));
  }
  for (int bps : {4, 8, 16}) {
    int blocks = sms * bps;
    printf("-- fq_mul: %d blocks/SM of 128 threads (ops = wide-MACs, 300 per product)\n", bps);
    report("fq_mul x1", (double)blocks * 128 * 1000 * 300, timeit([&] { t_fq_mul<1><<<blocks, 128>>>(io, 1000); }));
    report("fq_mul x2", (double)blocks * 128 * 1000 * 600, timeit([&] { t_fq_mul<2><<<blocks, 128>>>(io, 1000); }));
  }
  G1Affine* pts; cudaMalloc(&pts, 256 * sizeof(G1Affine));
  cudaMemcpy(pts, io, 256 * sizeof(G1Affine), cudaMemcpyDeviceToDevice);
  G1Xyzz* o2; cudaMalloc(&o2, sizeof(G1Xyzz));
  for (int bps : {2, 3, 4}) {
    int blocks = sms * bps;
    printf("-- madd: %d blocks/SM of 128 threads (ops = wide-MACs, 3000 per addition)\n", bps);
    report("madd", (double)blocks * 128 * 200 * 3000, timeit([&] { t_madd<<<blocks, 128>>>(pts, o2, 200); }));
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("status %s\n", cudaGetErrorString(e));
  return 0;
}
