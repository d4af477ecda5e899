// K3' — the G2 half of SetupArtifact: item i carries [s^i]G2 next to [s^i]G1
// (src/trusted_setup.rs:31-35, 64-72: one 256-bit blst_p2_mult on the G2 generator per item; it dominates the
// reference's setup time).  Same structure as the G1 kernel: the running power s^i comes from k_powers'
// chunked prefix products, the scalar multiplication becomes a fixed-base table walk — 32 unsigned 8-bit
// windows, table T2[j][d-1] = d * 2^(8 j) * G2 (32 x 255 affine points over Fq2, 1.5 MiB, L2-resident) — and
// each point leaves the kernel in its canonical form: the 96-byte compressed encoding of Serialize for
// G2Point (src/curves.rs:260-271).
#include "engine.cuh"
#include "g2.cuh"

namespace kzg {

constexpr int G2T_C = 8;
constexpr int G2T_W = 32;
constexpr int G2T_N = 255;          // multiples per window

__device__ __forceinline__ void inv_dev(Fq2& o, const Fq2& a) {
  Fq n, t;
  sqr(n, a.c0); sqr(t, a.c1); add(n, n, t);
  inv_gcd(t, n);          // uniform control flow (field.cuh), ~12x fewer instructions than Fermat
  mul(o.c0, a.c0, t);
  mul(t, a.c1, t);
  neg(o.c1, t);
}

// T2[j*255 + d-1] = d * 2^(8j) * G2 : thread = one table entry
__global__ void __launch_bounds__(64)
k_g2_table(G2Affine* __restrict__ tab) {
  uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= G2T_W * G2T_N) return;
  uint32_t j = t / G2T_N, d = t % G2T_N + 1;
  uint32_t k[9];
#pragma unroll
  for (int i = 0; i < 9; i++) k[i] = 0;
  uint32_t bit = j * G2T_C;
  k[bit >> 5] |= d << (bit & 31);          // 8-bit windows never straddle a limb
  G2Affine g; g2_generator(g);
  G2Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (int i = (int)bit + 7; i >= 0; i--) {
    dbl(acc);
    if ((k[i >> 5] >> (i & 31)) & 1) madd(acc, g.x, g.y);
  }
  G2Affine a;
  if (is_inf(acc)) set_inf(a);
  else { Fq2 w; inv_dev(w, acc.zzz); to_affine_with_inv(a, acc, w); }
  tab[t] = a;
}

__global__ void __launch_bounds__(64)
k_setup_points_g2(const Fr* __restrict__ scalars, uint64_t n, const G2Affine* __restrict__ tab,
                  uint8_t* __restrict__ out_c96) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Fr k = scalars[i];
  G2Xyzz acc; set_inf(acc);
#pragma unroll 1
  for (int j = 0; j < G2T_W; j++) {
    uint32_t d = (k.v[j >> 2] >> ((j & 3) * 8)) & 0xFFu;
    if (d != 0) {
      G2Affine p = tab[(size_t)j * G2T_N + (d - 1)];
      madd(acc, p.x, p.y);
    }
  }
  G2Affine a;
  if (is_inf(acc)) set_inf(a);
  else { Fq2 w; inv_dev(w, acc.zzz); to_affine_with_inv(a, acc, w); }
  compress(out_c96 + i * 96, a);
}

__global__ void k_powers_g2(Fr s_mont, uint64_t first, uint64_t n, uint32_t K, Fr* __restrict__ out) {
  uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t i0 = t * K;
  if (i0 >= n) return;
  Fr cur; pow_u64(cur, s_mont, first + i0);
  Fr c; from_mont(c, cur);
  for (uint32_t k = 0; k < K && i0 + k < n; k++) {
    out[i0 + k] = c;
    Fr nx; mul(nx, s_mont, c);
    c = nx;
  }
}

static inline unsigned blocks_for(uint64_t n, unsigned bs) { return (unsigned)((n + bs - 1) / bs); }

int32_t setup_generate_g2_dev(kzg_ctx* ctx, const uint8_t secret_be[32], uint64_t first, uint64_t n,
                              uint8_t* out_c96_dev) {
  if (n == 0) return KZG_OK;
  if (!ctx->g2table) {
    G2Affine* tab = nullptr;
    KZG_CUDA_OK(ctx, cudaMalloc(&tab, sizeof(G2Affine) * G2T_W * G2T_N));
    k_g2_table<<<blocks_for(G2T_W * G2T_N, 64), 64, 0, ctx->stream>>>(tab);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { cudaFree(tab); ctx->err = std::string("k_g2_table: ") + cudaGetErrorString(e); return KZG_E_CUDA; }
    ctx->g2table = tab;
  }
  uint8_t le[32];
  for (int i = 0; i < 32; i++) le[i] = secret_be[31 - i];
  Fr raw, s_mont; fr_load_le(raw, le); to_mont(s_mont, raw);
  void* pw;
  KZG_TRY(scratch_get(ctx, SCR_FR_B, n * sizeof(Fr), &pw));
  const uint32_t K = 32;
  k_powers_g2<<<blocks_for((n + K - 1) / K, 128), 128, 0, ctx->stream>>>(s_mont, first, n, K, reinterpret_cast<Fr*>(pw));
  KZG_LAUNCHED(ctx);
  k_setup_points_g2<<<blocks_for(n, 64), 64, 0, ctx->stream>>>(reinterpret_cast<Fr*>(pw), n,
                                                               reinterpret_cast<const G2Affine*>(ctx->g2table), out_c96_dev);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

}  // namespace kzg
