// Do the integer-multiply (FMA-heavy) pipe and the ALU pipe overlap on sm_100a?  NW carry-free mad.wide.u32 and NA
// two-limb add.cc/addc chains per inner iteration, all independent.  If time(NW, NA) ~ max(time(NW,0), time(0,NA)) the
// pipes overlap and an Fq product built from carry-free IMAD.WIDE + ALU carry handling could beat the carry-chained one.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench4.bin tools/ubench/ubench4.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048

template <int NW, int NA, bool CHAIN>
__global__ void __launch_bounds__(256) t_mix(uint32_t* out, uint32_t seed) {
  unsigned long long acc[NW ? NW : 1];
  uint32_t a[NW ? NW : 1];
  uint32_t xl[NA ? NA : 1], xh[NA ? NA : 1];
  uint32_t b = seed + blockIdx.x, y = seed * 3 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < (NW ? NW : 1); k++) { acc[k] = threadIdx.x + k; a[k] = seed * (k + 3) + threadIdx.x; }
#pragma unroll
  for (int k = 0; k < (NA ? NA : 1); k++) { xl[k] = threadIdx.x * 5 + k; xh[k] = k; }
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if constexpr (CHAIN && NW > 0) {
        // one carry chain through the NW accumulators (lo/hi pairs), like a row of the shipped Fq product
        uint32_t lo[NW ? NW : 1], hi[NW ? NW : 1];
#pragma unroll
        for (int k = 0; k < NW; k++) { lo[k] = (uint32_t)acc[k]; hi[k] = (uint32_t)(acc[k] >> 32); }
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(a[0]), "r"(b));
#pragma unroll
        for (int k = 1; k < NW; k++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(a[k]), "r"(b));
#pragma unroll
        for (int k = 0; k < NW; k++) acc[k] = ((unsigned long long)hi[k] << 32) | lo[k];
      } else {
#pragma unroll
        for (int k = 0; k < NW; k++)
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(b));
      }
#pragma unroll
      for (int k = 0; k < NA; k++)
        asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(xl[k]), "+r"(xh[k]) : "r"(y), "r"(b));
      b += 7;       // a new multiplier every round: ptxas must not reuse a product
    }
  }
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < (NW ? NW : 1); k++) s ^= acc[k];
#pragma unroll
  for (int k = 0; k < (NA ? NA : 1); k++) s ^= xl[k] ^ xh[k];
  if (s == 0x12345678ull) out[0] = 1;
}

template <int NW, int NA, bool CHAIN>
static void run(const char* name, int sms, int blocks_per_sm) {
  uint32_t* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  int grid = sms * blocks_per_sm;
  t_mix<NW, NA, CHAIN><<<grid, 256>>>(out, 12345u);
  cudaDeviceSynchronize();
  cudaEventRecord(e0);
  t_mix<NW, NA, CHAIN><<<grid, 256>>>(out, 12345u);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  // warp-iterations per SMSP: (blocks_per_sm * 8 warps / 4 SMSP) * ITERS * 4
  double warp_iters = (double)blocks_per_sm * 2 * ITERS * 4;
  double cyc = ms * 1e-3 * clk_khz * 1e3 / warp_iters;
  printf("%-44s %2d blk/SM  %8.3f ms  %7.2f cyc per (NW=%d wide + NA=%d add pairs) per SMSP  [%.2f per wide]\n", name, blocks_per_sm, ms, cyc, NW, NA,
         NW ? cyc / NW : 0.0);
  cudaFree(out);
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  printf("SMs %d\n", sms);
  for (int b : {4, 8}) {
    run<8, 0, false>("wide x8", sms, b);
    run<0, 8, false>("add pairs x8", sms, b);
    run<0, 16, false>("add pairs x16", sms, b);
    run<8, 4, false>("wide x8 + add pairs x4", sms, b);
    run<8, 8, false>("wide x8 + add pairs x8", sms, b);
    run<8, 16, false>("wide x8 + add pairs x16", sms, b);
    run<6, 0, true>("wide carry chain of 6", sms, b);
    run<6, 6, true>("wide carry chain of 6 + add pairs x6", sms, b);
    run<6, 12, true>("wide carry chain of 6 + add pairs x12", sms, b);
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
