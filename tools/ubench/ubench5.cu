// Round 2: what IS the integer-multiply roofline of sm_100a?  Pure dependent chains (8 per thread, 16 warps per
// scheduler: throughput-, not latency-bound), every loop body checked with cuobjdump to contain ONLY the instruction
// under test plus register moves:
//   imad      p <- p * p + p            (32-bit)      IMAD            nominal 64 lanes / clk / SM
//   imad_hi   p <- hi(p * q)            (32-bit)      IMAD.HI.U32
//   wide      p <- lo(p) * hi(p)        (32x32->64)   IMAD.WIDE.U32 , RZ addend
//   wide_acc  p <- lo(p) * hi(p) + p    (32x32+64)    IMAD.WIDE.U32 with a 64-bit addend
//   chain6    one 6-limb row a_i * b with carries    IMAD.WIDE.U32(.X) carry chain, as in the shipped Fq product
// Result on B200 (profiles/ubench_imad_r02.txt): a 32x32->64 product issues at HALF the 32-bit IMAD rate — 32 wide
// multiplies / clk / SM, 9.3e12 / s at 1965 MHz — so the "64 lanes" of SURVEY 8d are a peak for 32-bit products only,
// and the shipped Fq product (300 wide-MAC, kzg_probe_fq_mul) already runs at ~98 % of the wide-multiply issue rate.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/ubench5.bin tools/ubench/ubench5.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 2048
#define CH 8

template <int MODE>
__global__ void __launch_bounds__(256) t_pipe(uint32_t* out, uint32_t seed) {
  uint64_t p[CH];
  uint32_t q = (seed ^ 0x9E3779B9u) + blockIdx.x * 977u + threadIdx.x;
#pragma unroll
  for (int k = 0; k < CH; k++) p[k] = ((uint64_t)(seed + threadIdx.x * 2654435761u + k * 0x85EBCA6Bu) << 32) | (q + k * 3u);
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int r = 0; r < 4; r++) {
      if constexpr (MODE == 4) {
        uint32_t lo[CH], hi[CH];
#pragma unroll
        for (int k = 0; k < CH; k++) { lo[k] = (uint32_t)p[k]; hi[k] = (uint32_t)(p[k] >> 32); }
        // row of 6 (+2 idle): lo/hi pairs accumulate a_k * q with the carry running through the row
        asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[0]), "+r"(hi[0]) : "r"(hi[1]), "r"(q));
#pragma unroll
        for (int k = 1; k < CH; k++)
          asm volatile("madc.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(lo[k]), "+r"(hi[k]) : "r"(hi[(k + 1) % CH]), "r"(q));
#pragma unroll
        for (int k = 0; k < CH; k++) p[k] = ((uint64_t)hi[k] << 32) | lo[k];
      } else {
#pragma unroll
        for (int k = 0; k < CH; k++) {
          uint32_t lo = (uint32_t)p[k], hi = (uint32_t)(p[k] >> 32);
          if constexpr (MODE == 0) { asm volatile("mad.lo.u32 %0, %0, %0, %0;" : "+r"(lo)); p[k] = ((uint64_t)hi << 32) | lo; }
          if constexpr (MODE == 1) { asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(lo) : "r"(q)); p[k] = ((uint64_t)hi << 32) | lo; }
          if constexpr (MODE == 2) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p[k]) : "r"(lo), "r"(hi));
          if constexpr (MODE == 3) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(p[k]) : "r"(lo), "r"(hi));
        }
      }
    }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < CH; k++) s ^= p[k];
  if (s == 0x1234567ull) out[0] = (uint32_t)s;
}

template <int MODE>
static void run(const char* name, int sms, double ghz) {
  uint32_t* out; cudaMalloc(&out, 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int grid = sms * 8;
  t_pipe<MODE><<<grid, 256>>>(out, 12345u);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    t_pipe<MODE><<<grid, 256>>>(out, 777u + rep);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  const double ops = (double)grid * 256 * ITERS * 4.0 * CH;
  const double rate = ops / (best * 1e-3);
  printf("%-10s %.3e ops/s  %6.2f lanes/clk/SM  (%.2f cyc per warp-instr per SMSP)\n", name, rate, rate / (sms * ghz * 1e9),
         32.0 / (rate / (sms * ghz * 1e9) / 4.0));
  cudaFree(out);
}

int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  const double ghz = khz / 1e6;
  printf("%s, %d SMs, %.3f GHz max SM clock; 8 blocks of 256 threads per SM, 8 chains per thread\n", prop.name, prop.multiProcessorCount, ghz);
  run<0>("imad", prop.multiProcessorCount, ghz);
  run<1>("imad_hi", prop.multiProcessorCount, ghz);
  run<2>("wide", prop.multiProcessorCount, ghz);
  run<3>("wide_acc", prop.multiProcessorCount, ghz);
  run<4>("chain8", prop.multiProcessorCount, ghz);
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
