// Reduced-radix (13 x 30-bit) Montgomery product for Fq: carry-free IMAD.WIDE column accumulation.
// Question: does it beat the shipped 12 x 32-bit carry-chained product (fq_mul_ptx) on sm_100a?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/ubench3 tools/ubench/ubench3.cu && gpurun_out/ubench3
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include <vector>
#include "../../kzg_poly_commit_exploration_b200/csrc/g1.cuh"
using namespace kzg;

constexpr uint32_t M30 = (1u << 30) - 1;
struct P30 { uint32_t v[13]; };
__constant__ uint32_t c_p30[13];
__constant__ uint32_t c_inv30;

template <bool SQR>
__device__ __forceinline__ void mont30(uint32_t* o, const uint32_t* a, const uint32_t* b, const uint32_t* p, uint32_t inv) {
  uint64_t c[26];
#pragma unroll
  for (int k = 0; k < 26; k++) c[k] = 0;
  if (!SQR) {
#pragma unroll
    for (int i = 0; i < 13; i++)
#pragma unroll
      for (int j = 0; j < 13; j++) c[i + j] += (uint64_t)a[i] * b[j];
  } else {
    uint32_t a2[13];
#pragma unroll
    for (int i = 0; i < 13; i++) a2[i] = a[i] << 1;
#pragma unroll
    for (int i = 0; i < 13; i++) {
      c[2 * i] += (uint64_t)a[i] * a[i];
#pragma unroll
      for (int j = i + 1; j < 13; j++) c[i + j] += (uint64_t)a2[i] * a[j];
    }
  }
#pragma unroll
  for (int k = 0; k < 25; k++) { c[k + 1] += c[k] >> 30; c[k] &= M30; }
#pragma unroll
  for (int i = 0; i < 13; i++) {
    uint32_t m = ((uint32_t)c[i] * inv) & M30;
#pragma unroll
    for (int j = 0; j < 13; j++) c[i + j] += (uint64_t)m * p[j];
    c[i + 1] += c[i] >> 30;
  }
#pragma unroll
  for (int k = 13; k < 25; k++) { c[k + 1] += c[k] >> 30; c[k] &= M30; }
#pragma unroll
  for (int k = 0; k < 13; k++) o[k] = (uint32_t)c[13 + k];
}

template <int MODE>
__global__ void __launch_bounds__(128) t_mont30(uint32_t* io, int iters) {
  uint32_t a[13], b[13], p[13];
  uint32_t inv = c_inv30;
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 13; k++) { a[k] = io[t * 13 + k] & M30; b[k] = (a[k] * 2654435761u + k) & M30; p[k] = c_p30[k]; }
#pragma unroll 1
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) mont30<false>(a, a, b, p, inv);
    else if (MODE == 1) mont30<true>(a, a, a, p, inv);
    else { mont30<false>(a, a, b, c_p30, inv); }
  }
#pragma unroll
  for (int k = 0; k < 13; k++) io[t * 13 + k] = a[k];
}

__global__ void __launch_bounds__(128) t_fq_mul(Fq* io, int iters) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fq a = io[t], b = a; b.v[0] ^= 0x55;
#pragma unroll 1
  for (int it = 0; it < iters; it++) mul(a, a, b);
  io[t] = a;
}
__global__ void __launch_bounds__(128) t_fq_sqr(Fq* io, int iters) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  Fq a = io[t];
#pragma unroll 1
  for (int it = 0; it < iters; it++) sqr(a, a);
  io[t] = a;
}

typedef unsigned __int128 u128;
// host check: big integer as 13x30 limbs -> compare a*b*R^-1 mod q by a slow bit-serial reference
static void host_mont30(uint32_t* o, const uint32_t* a, const uint32_t* b, const uint32_t* p, uint32_t inv) {
  uint64_t c[27] = {0};
  for (int i = 0; i < 13; i++) for (int j = 0; j < 13; j++) {
    u128 s = (u128)c[i + j] + (u128)a[i] * b[j];
    if (s >> 64) { printf("overflow product phase\n"); }
    c[i + j] = (uint64_t)s;
  }
  for (int k = 0; k < 25; k++) { c[k + 1] += c[k] >> 30; c[k] &= M30; }
  for (int i = 0; i < 13; i++) {
    uint32_t m = ((uint32_t)c[i] * inv) & M30;
    for (int j = 0; j < 13; j++) {
      u128 s = (u128)c[i + j] + (u128)m * p[j];
      if (s >> 64) { printf("overflow reduce phase\n"); }
      c[i + j] = (uint64_t)s;
    }
    if (c[i] & M30) printf("low bits not cleared\n");
    c[i + 1] += c[i] >> 30;
  }
  for (int k = 13; k < 25; k++) { c[k + 1] += c[k] >> 30; c[k] &= M30; }
  for (int k = 0; k < 13; k++) o[k] = (uint32_t)c[13 + k];
}

int main() {
  // q in 30-bit limbs
  uint32_t q32[12];
  for (int i = 0; i < 12; i++) q32[i] = h_FQ_P[i];
  uint32_t p30[13];
  for (int k = 0; k < 13; k++) {
    int bit = 30 * k; uint64_t w = 0;
    int wi = bit >> 5, sh = bit & 31;
    w = q32[wi]; if (wi + 1 < 12) w |= (uint64_t)q32[wi + 1] << 32;
    if (wi >= 12) w = 0;
    p30[k] = (uint32_t)(w >> sh) & M30;
  }
  uint32_t inv = 1;  // -q^-1 mod 2^30
  for (int i = 0; i < 6; i++) inv *= 2 - p30[0] * inv;
  inv = (0u - inv) & M30;
  if (((p30[0] * inv + 1) & M30) != 0) { printf("bad inv\n"); return 1; }
  cudaMemcpyToSymbol(c_p30, p30, sizeof(p30));
  cudaMemcpyToSymbol(c_inv30, &inv, 4);
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  int sms = prop.multiProcessorCount;
  double clk = 1.965e9;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  auto timeit = [&](auto f) { f(); cudaDeviceSynchronize(); float best = 1e30f; for (int r = 0; r < 3; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; } return best; };
  for (int bps : {3, 4, 6, 8}) {
    int blocks = sms * bps; size_t nt = (size_t)blocks * 128;
    uint32_t* io; cudaMalloc(&io, nt * 13 * 4);
    std::vector<uint32_t> h(nt * 13);
    for (size_t i = 0; i < h.size(); i++) h[i] = (uint32_t)(i * 2654435761u + 12345) & M30;
    for (size_t i = 0; i < nt; i++) h[i * 13 + 12] &= 0x1fff;  // < 2^373
    Fq* io2; cudaMalloc(&io2, nt * sizeof(Fq)); cudaMemset(io2, 0x17, nt * sizeof(Fq));
    printf("-- %d blocks/SM of 128 threads\n", bps);
    auto rep = [&](const char* name, float ms) {
      double prods = (double)nt * 1000 / (ms * 1e-3);
      printf("%-28s %.3e products/s  %.1f cyc per warp-product per SMSP\n", name, prods, 4.0 * 32.0 * sms * clk / prods / 1.0);
    };
    cudaMemcpy(io, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    // correctness of one iteration
    t_mont30<0><<<blocks, 128>>>(io, 1);
    std::vector<uint32_t> g(h.size()); cudaMemcpy(g.data(), io, h.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (size_t t = 0; t < 64; t++) {
      uint32_t a[13], b[13], o[13];
      for (int k = 0; k < 13; k++) { a[k] = h[t * 13 + k] & M30; b[k] = (a[k] * 2654435761u + k) & M30; }
      host_mont30(o, a, b, p30, inv);
      for (int k = 0; k < 13; k++) if (o[k] != g[t * 13 + k]) bad++;
    }
    printf("check vs host: %s\n", bad ? "MISMATCH" : "ok");
    rep("mont30 mul (p in regs)", timeit([&] { t_mont30<0><<<blocks, 128>>>(io, 1000); }));
    rep("mont30 sqr (p in regs)", timeit([&] { t_mont30<1><<<blocks, 128>>>(io, 1000); }));
    rep("mont30 mul (p const mem)", timeit([&] { t_mont30<2><<<blocks, 128>>>(io, 1000); }));
    rep("fq_mul_ptx 12x32", timeit([&] { t_fq_mul<<<blocks, 128>>>(io2, 1000); }));
    rep("fq_sqr_ptx 12x32", timeit([&] { t_fq_sqr<<<blocks, 128>>>(io2, 1000); }));
    cudaFree(io); cudaFree(io2);
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
