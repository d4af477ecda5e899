// K4 — Horner evaluation and the (p(x) - p(z)) / (x - z) quotient as ONE affine-recurrence scan.
//
// Replaces Polynomial::evaluate (src/polynomial.rs:112-123, one bespoke pow per term,
// src/scalar.rs:122-187), Polynomial::sub (:128-145) and Polynomial::divide_by_root (:150-195).
// The synthetic-division intermediates ARE Horner's intermediates:
//   t_{n-1} = c_{n-1};  t_i = c_i + z * t_{i+1};  q_i = t_{i+1};  y = p(z) = t_0,
// so evaluation and quotient are one pass: read n x 32 B, write (n-1) x 32 B.
//
// Parallel form: a thread owns L = 8 consecutive coefficients (a map t -> z^L t + h), a block
// of 512 threads owns a 4096-coefficient segment and suffix-scans the maps in shared memory
// (all multipliers are powers of z, so only the h component is scanned); segments of a long
// polynomial are chained by a tiny carry kernel.  HBM-bound in the blob batch (64 B per
// coefficient), latency-bound for a single polynomial.
//
// Arithmetic trick: with zm = z*R (Montgomery form) and t canonical, mont_mul(zm, t) = z*t
// canonical — so coefficients never need converting to or from Montgomery form.
#include "engine.cuh"

namespace kzg {

constexpr int POLY_L = 8;
constexpr int POLY_THREADS = 512;
constexpr int POLY_SEG = POLY_L * POLY_THREADS;  // 4096 coefficients per block

__device__ __forceinline__ void ld_fr(Fr& o, const uint8_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  o.v[0] = a.x; o.v[1] = a.y; o.v[2] = a.z; o.v[3] = a.w;
  o.v[4] = b.x; o.v[5] = b.y; o.v[6] = b.z; o.v[7] = b.w;
}
__device__ __forceinline__ void st_fr(uint8_t* p, const Fr& a) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(a.v[0], a.v[1], a.v[2], a.v[3]);
  q[1] = make_uint4(a.v[4], a.v[5], a.v[6], a.v[7]);
}

// mode 0: write only the segment value H[blob][seg] (zero carry-in)
// mode 1: full pass with carry-in tin[blob][seg] (or 0 when tin == nullptr): writes y and q
template <int MODE>
__global__ void __launch_bounds__(POLY_THREADS)
k_poly_scan(const uint8_t* __restrict__ coeffs, uint64_t n, const uint8_t* __restrict__ z_all,
            const Fr* __restrict__ tin, Fr* __restrict__ hseg, uint8_t* __restrict__ y_out,
            uint8_t* __restrict__ q_out, uint32_t nseg) {
  __shared__ Fr sh[POLY_THREADS];
  const uint32_t seg = blockIdx.x, blob = blockIdx.y, k = threadIdx.x;
  const uint8_t* cb = coeffs + (size_t)blob * n * 32;
  const uint64_t base = (uint64_t)seg * POLY_SEG + (uint64_t)k * POLY_L;

  Fr zm;
  { Fr zr; ld_fr(zr, z_all + (size_t)blob * 32); to_mont(zm, zr); }

  Fr c[POLY_L];
#pragma unroll
  for (int j = 0; j < POLY_L; j++) {
    if (base + j < n) { ld_fr(c[j], cb + (base + j) * 32); fr_canon_reduce(c[j]); }
    else set_zero(c[j]);
  }
  // local Horner (zero carry-in)
  Fr h = c[POLY_L - 1];
#pragma unroll
  for (int j = POLY_L - 2; j >= 0; j--) { Fr t; mul(t, zm, h); add(h, c[j], t); }

  Fr pw = zm;  // z^L, Montgomery
#pragma unroll
  for (int s = 1; s < POLY_L; s <<= 1) sqr(pw, pw);

  Fr carry; set_zero(carry);
  if (MODE == 1 && tin != nullptr) carry = tin[(size_t)blob * nseg + seg];
  if (k == POLY_THREADS - 1 && MODE == 1 && tin != nullptr) { Fr t; mul(t, pw, carry); add(h, h, t); }

  // suffix scan: V_k = sum_j h_{k+j} z^{L j}
  Fr V = h;
#pragma unroll 1
  for (int s = 1; s < POLY_THREADS; s <<= 1) {
    sh[k] = V;
    __syncthreads();
    if (k + s < POLY_THREADS) { Fr t; mul(t, pw, sh[k + s]); add(V, V, t); }
    __syncthreads();
    sqr(pw, pw);
  }
  if (MODE == 0) {
    if (k == 0) hseg[(size_t)blob * nseg + seg] = V;
    return;
  }
  sh[k] = V;
  __syncthreads();
  Fr t = (k + 1 < POLY_THREADS) ? sh[k + 1] : carry;   // t_{base + L}
  uint8_t* qb = q_out ? q_out + (size_t)blob * n * 32 : nullptr;
#pragma unroll
  for (int j = POLY_L - 1; j >= 0; j--) {
    Fr u; mul(u, zm, t); add(t, c[j], u);               // t_{base + j}
    uint64_t idx = base + j;
    if (idx < n) {
      if (idx == 0) st_fr(y_out + (size_t)blob * 32, t);
      else if (qb) st_fr(qb + (idx - 1) * 32, t);
      if (qb && idx == n - 1) { Fr zero; set_zero(zero); st_fr(qb + idx * 32, zero); }
    }
  }
}

// chain the segments of each polynomial: tin[s] = H[s+1] + z^SEG * tin[s+1]
__global__ void k_poly_carry(const Fr* __restrict__ hseg, const uint8_t* __restrict__ z_all,
                             Fr* __restrict__ tin, uint32_t nseg, uint32_t blobs) {
  uint32_t blob = blockIdx.x * blockDim.x + threadIdx.x;
  if (blob >= blobs) return;
  Fr zm;
  { Fr zr; ld_fr(zr, z_all + (size_t)blob * 32); to_mont(zm, zr); }
  Fr pw = zm;
  for (int s = 1; s < POLY_SEG; s <<= 1) sqr(pw, pw);
  Fr t; set_zero(t);
  tin[(size_t)blob * nseg + nseg - 1] = t;
  for (int s = (int)nseg - 2; s >= 0; s--) {
    Fr u; mul(u, pw, t);
    add(t, hseg[(size_t)blob * nseg + s + 1], u);
    tin[(size_t)blob * nseg + s] = t;
  }
}

__global__ void k_zero32(uint8_t* y, uint32_t blobs) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < blobs * 8) reinterpret_cast<uint32_t*>(y)[i] = 0;
}

// Polynomial::try_from trims trailing zeros (src/polynomial.rs:63-72) and commit compares the TRIMMED length with the
// setup (:201-205): rows of a device-resident batch that are longer than the setup are fine when everything from
// index `limit` on is zero mod r.  One pass over the tail columns, result through a flag word.
__global__ void k_any_nonzero_beyond(const uint8_t* __restrict__ coeffs, uint64_t blobs, uint64_t n, uint64_t limit,
                                     uint32_t* __restrict__ flag) {
  const uint64_t tail = n - limit;
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= blobs * tail) return;
  const uint64_t row = i / tail, col = limit + i % tail;
  Fr c; ld_fr(c, coeffs + (row * n + col) * 32);
  fr_canon_reduce(c);
  if (!is_zero(c)) *flag = 1u;
}

int32_t any_nonzero_beyond(kzg_ctx* ctx, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n, uint64_t limit, bool* any) {
  *any = false;
  if (blobs == 0 || limit >= n) return KZG_OK;
  void* m;
  KZG_TRY(scratch_get(ctx, SCR_MISC, 1024, &m));
  uint32_t* flag = reinterpret_cast<uint32_t*>(m) + 200;       // clear of the 192-byte result slot other calls keep here
  KZG_CUDA_OK(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
  const uint64_t total = blobs * (n - limit);
  k_any_nonzero_beyond<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(coeffs_dev, blobs, n, limit, flag);
  KZG_LAUNCHED(ctx);
  uint32_t h = 0;
  KZG_CUDA_OK(ctx, cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
  KZG_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  *any = h != 0;
  return KZG_OK;
}

// one launch set covers at most 65535 polynomials (grid.y); larger batches run in row chunks
static int32_t poly_eval_quotient_rows(kzg_ctx* ctx, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n,
                                       const uint8_t* z_dev, uint8_t* y_dev, uint8_t* q_dev) {
  uint32_t nseg = (uint32_t)((n + POLY_SEG - 1) / POLY_SEG);
  dim3 grid(nseg, (unsigned)blobs);
  Fr* tin = nullptr;
  if (nseg > 1) {
    void* p;
    KZG_TRY(scratch_get(ctx, SCR_SEG, (size_t)blobs * nseg * sizeof(Fr) * 2, &p));
    Fr* hseg = reinterpret_cast<Fr*>(p);
    tin = hseg + (size_t)blobs * nseg;
    k_poly_scan<0><<<grid, POLY_THREADS, 0, ctx->stream>>>(coeffs_dev, n, z_dev, nullptr, hseg, nullptr, nullptr, nseg);
    KZG_LAUNCHED(ctx);
    k_poly_carry<<<(unsigned)((blobs + 63) / 64), 64, 0, ctx->stream>>>(hseg, z_dev, tin, nseg, (uint32_t)blobs);
    KZG_LAUNCHED(ctx);
  }
  k_poly_scan<1><<<grid, POLY_THREADS, 0, ctx->stream>>>(coeffs_dev, n, z_dev, tin, nullptr, y_dev, q_dev, nseg);
  KZG_LAUNCHED(ctx);
  return KZG_OK;
}

int32_t poly_eval_quotient(kzg_ctx* ctx, const uint8_t* coeffs_dev, uint64_t blobs, uint64_t n,
                           const uint8_t* z_dev, uint8_t* y_dev, uint8_t* q_dev) {
  if (blobs == 0) return KZG_OK;
  if (n == 0) {  // empty polynomial evaluates to 0 (src/polynomial.rs:113)
    k_zero32<<<(unsigned)((blobs * 8 + 255) / 256), 256, 0, ctx->stream>>>(y_dev, (uint32_t)blobs);
    KZG_LAUNCHED(ctx);
    return KZG_OK;
  }
  constexpr uint64_t MAX_ROWS = 65535;            // grid.y
  for (uint64_t r0 = 0; r0 < blobs; r0 += MAX_ROWS) {
    const uint64_t rows = blobs - r0 < MAX_ROWS ? blobs - r0 : MAX_ROWS;
    KZG_TRY(poly_eval_quotient_rows(ctx, coeffs_dev + r0 * n * 32, rows, n, z_dev + r0 * 32, y_dev + r0 * 32,
                                    q_dev ? q_dev + r0 * n * 32 : nullptr));
  }
  return KZG_OK;
}

}  // namespace kzg
