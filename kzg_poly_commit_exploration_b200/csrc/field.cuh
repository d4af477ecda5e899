// Fq (381-bit, 12x32) and Fr (255-bit, 8x32) Montgomery arithmetic.
//
// Replaces, call site by call site, what the reference gets from blst:
//   blst_fr_mul/_add/_sub/_cneg            src/scalar.rs:114,195,206,215
//   blst_fr_from_hexascii (bytes -> Fr)    src/scalar.rs:34,58,70
//   blst_scalar_from_fr + *endian_from_*   src/scalar.rs:86-104
//   and the Fp arithmetic inside every blst_p1_* (src/curves.rs, src/trusted_setup.rs).
//
// Device code runs the generated straight-line PTX (field_ptx_gen.cuh: IMAD.WIDE.U32.X
// carry chains).  Host code runs a portable 64-bit CIOS with identical results; it backs
// the single-element C-ABI helpers (Scalar::mul, G1Point::add ...) and lets every
// formula above this layer be exercised by the CPU test-suite with the same source.
#pragma once
#include <cstdint>
#include <cstring>
#include "consts_gen.cuh"
#include "field_ptx_gen.cuh"

#if defined(__CUDACC__)
#define KZG_HD __host__ __device__ __forceinline__
#define KZG_D __device__ __forceinline__
#else
#define KZG_HD inline
#define KZG_D inline
#endif

namespace kzg {

// ---- constants: __constant__ for device, plain arrays for host -----------------------
#if defined(__CUDACC__)
#define KZG_DEF_CONST(name, n, ...) \
  __device__ __constant__ const uint32_t d_##name[n] = __VA_ARGS__; \
  static const uint32_t h_##name[n] = __VA_ARGS__;
#else
#define KZG_DEF_CONST(name, n, ...) static const uint32_t h_##name[n] = __VA_ARGS__;
#endif
KZG_CONST_LIST(KZG_DEF_CONST)
#undef KZG_DEF_CONST

#if defined(__CUDA_ARCH__)
#define KZG_C(name) d_##name
#else
#define KZG_C(name) h_##name
#endif

struct alignas(16) Fq { uint32_t v[12]; };
struct alignas(16) Fr { uint32_t v[8]; };

// ---- host reference arithmetic (portable, 64-bit accumulators) -----------------------
template <int N>
inline void host_mont_mul(uint32_t* o, const uint32_t* a, const uint32_t* b, const uint32_t* p, uint32_t inv) {
  uint32_t t[N + 2];
  for (int i = 0; i < N + 2; i++) t[i] = 0;
  for (int i = 0; i < N; i++) {
    uint64_t c = 0;
    for (int j = 0; j < N; j++) {
      uint64_t s = (uint64_t)a[j] * b[i] + t[j] + c;
      t[j] = (uint32_t)s; c = s >> 32;
    }
    uint64_t s = (uint64_t)t[N] + c;
    t[N] = (uint32_t)s; t[N + 1] = (uint32_t)(s >> 32);
    uint32_t m = t[0] * inv;
    c = ((uint64_t)m * p[0] + t[0]) >> 32;
    for (int j = 1; j < N; j++) {
      uint64_t s2 = (uint64_t)m * p[j] + t[j] + c;
      t[j - 1] = (uint32_t)s2; c = s2 >> 32;
    }
    s = (uint64_t)t[N] + c;
    t[N - 1] = (uint32_t)s;
    t[N] = t[N + 1] + (uint32_t)(s >> 32);
  }
  // conditional subtract
  uint32_t u[N]; uint64_t br = 0;
  for (int j = 0; j < N; j++) {
    uint64_t d = (uint64_t)t[j] - p[j] - br;
    u[j] = (uint32_t)d; br = (d >> 63) & 1;
  }
  bool ge = t[N] != 0 || br == 0;
  for (int j = 0; j < N; j++) o[j] = ge ? u[j] : t[j];
}

template <int N>
inline void host_add_mod(uint32_t* o, const uint32_t* a, const uint32_t* b, const uint32_t* p) {
  uint32_t t[N], u[N]; uint64_t c = 0, br = 0;
  for (int j = 0; j < N; j++) { uint64_t s = (uint64_t)a[j] + b[j] + c; t[j] = (uint32_t)s; c = s >> 32; }
  for (int j = 0; j < N; j++) { uint64_t d = (uint64_t)t[j] - p[j] - br; u[j] = (uint32_t)d; br = (d >> 63) & 1; }
  bool ge = c != 0 || br == 0;
  for (int j = 0; j < N; j++) o[j] = ge ? u[j] : t[j];
}

template <int N>
inline void host_sub_mod(uint32_t* o, const uint32_t* a, const uint32_t* b, const uint32_t* p) {
  uint32_t t[N]; uint64_t br = 0, c = 0;
  for (int j = 0; j < N; j++) { uint64_t d = (uint64_t)a[j] - b[j] - br; t[j] = (uint32_t)d; br = (d >> 63) & 1; }
  for (int j = 0; j < N; j++) { uint64_t s = (uint64_t)t[j] + (br ? p[j] : 0) + c; o[j] = (uint32_t)s; c = s >> 32; }
}

// ---- dispatch ---------------------------------------------------------------------
#if defined(__CUDACC__) && defined(KZG_FQ_NOINLINE)
// one out-of-line copy of the product per translation unit (operands travel in registers): for the
// latency-bound kernels, where instruction fetch of the inlined 6 KB sequence is the bottleneck
static __device__ __noinline__ Fq fq_mul_call(Fq a, Fq b) {
  Fq o;
#if defined(__CUDA_ARCH__)
  fq_mul_ptx(o.v, a.v, b.v);
#endif
  return o;
}
#endif
KZG_HD void mul(Fq& o, const Fq& a, const Fq& b) {
#if defined(__CUDA_ARCH__) && defined(KZG_FQ_NOINLINE)
  o = fq_mul_call(a, b);
#elif defined(__CUDA_ARCH__)
  fq_mul_ptx(o.v, a.v, b.v);
#else
  host_mont_mul<12>(o.v, a.v, b.v, h_FQ_P, FQ_INV32);
#endif
}
KZG_HD void sqr(Fq& o, const Fq& a) {
#if defined(__CUDA_ARCH__) && defined(KZG_FQ_NOINLINE)
  o = fq_mul_call(a, a);
#elif defined(__CUDA_ARCH__)
  fq_sqr_ptx(o.v, a.v);
#else
  host_mont_mul<12>(o.v, a.v, a.v, h_FQ_P, FQ_INV32);
#endif
}
KZG_HD void add(Fq& o, const Fq& a, const Fq& b) {
#if defined(__CUDA_ARCH__)
  fq_add_ptx(o.v, a.v, b.v);
#else
  host_add_mod<12>(o.v, a.v, b.v, h_FQ_P);
#endif
}
KZG_HD void sub(Fq& o, const Fq& a, const Fq& b) {
#if defined(__CUDA_ARCH__)
  fq_sub_ptx(o.v, a.v, b.v);
#else
  host_sub_mod<12>(o.v, a.v, b.v, h_FQ_P);
#endif
}
KZG_HD void mul(Fr& o, const Fr& a, const Fr& b) {
#if defined(__CUDA_ARCH__)
  fr_mul_ptx(o.v, a.v, b.v);
#else
  host_mont_mul<8>(o.v, a.v, b.v, h_FR_P, FR_INV32);
#endif
}
KZG_HD void sqr(Fr& o, const Fr& a) { mul(o, a, a); }
KZG_HD void add(Fr& o, const Fr& a, const Fr& b) {
#if defined(__CUDA_ARCH__)
  fr_add_ptx(o.v, a.v, b.v);
#else
  host_add_mod<8>(o.v, a.v, b.v, h_FR_P);
#endif
}
KZG_HD void sub(Fr& o, const Fr& a, const Fr& b) {
#if defined(__CUDA_ARCH__)
  fr_sub_ptx(o.v, a.v, b.v);
#else
  host_sub_mod<8>(o.v, a.v, b.v, h_FR_P);
#endif
}

// ---- small helpers (both fields) ---------------------------------------------------
template <class F> KZG_HD constexpr int nlimbs() { return (int)(sizeof(F) / 4); }

template <class F> KZG_HD bool is_zero(const F& a) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < nlimbs<F>(); i++) acc |= a.v[i];
  return acc == 0;
}
template <class F> KZG_HD bool eq(const F& a, const F& b) {
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < nlimbs<F>(); i++) acc |= a.v[i] ^ b.v[i];
  return acc == 0;
}
template <class F> KZG_HD void set_zero(F& a) {
#pragma unroll
  for (int i = 0; i < nlimbs<F>(); i++) a.v[i] = 0;
}
KZG_HD void set_one(Fq& a) {
#pragma unroll
  for (int i = 0; i < 12; i++) a.v[i] = KZG_C(FQ_ONE)[i];
}
KZG_HD void set_one(Fr& a) {
#pragma unroll
  for (int i = 0; i < 8; i++) a.v[i] = KZG_C(FR_ONE)[i];
}
template <class F> KZG_HD void neg(F& o, const F& a) {
  F z; set_zero(z);
  sub(o, z, a);
}
template <class F> KZG_HD void dbl(F& o, const F& a) { add(o, a, a); }

// Montgomery <-> canonical
KZG_HD void to_mont(Fq& o, const Fq& a) {
  Fq r2;
#pragma unroll
  for (int i = 0; i < 12; i++) r2.v[i] = KZG_C(FQ_R2)[i];
  mul(o, r2, a);
}
KZG_HD void from_mont(Fq& o, const Fq& a) {
  Fq one; set_zero(one); one.v[0] = 1;
  mul(o, a, one);
}
// bytes -> Fr: any 256-bit integer (possibly >= r) * R^2 -> fully reduced Montgomery form;
// the same thing blst_fr_from_hexascii does for Scalar::from_le_bytes (src/scalar.rs:54-61).
KZG_HD void to_mont(Fr& o, const Fr& a) {
  Fr r2;
#pragma unroll
  for (int i = 0; i < 8; i++) r2.v[i] = KZG_C(FR_R2)[i];
  mul(o, r2, a);
}
KZG_HD void from_mont(Fr& o, const Fr& a) {
  Fr one; set_zero(one); one.v[0] = 1;
  mul(o, a, one);
}

// canonical reduction of a raw 256-bit integer to [0, r): 2^256 < 3r, so two conditional
// subtractions (Scalar::from_le_bytes semantics without the Montgomery conversion — the MSM
// needs canonical integers, src/curves.rs:90-96 feeds Scalar::to_le_bytes to blst_p1_mult).
KZG_HD void fr_canon_reduce(Fr& a) {
#pragma unroll
  for (int rep = 0; rep < 2; rep++) {
    uint32_t u[8]; uint32_t br = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint64_t d = (uint64_t)a.v[j] - KZG_C(FR_P)[j] - br;
      u[j] = (uint32_t)d; br = (uint32_t)(d >> 63);
    }
    if (!br) {
#pragma unroll
      for (int j = 0; j < 8; j++) a.v[j] = u[j];
    }
  }
}

// a > (q-1)/2 on the canonical value — decides the ZCash sign bit (src/curves.rs:106).
KZG_HD bool fq_canon_gt_half(const Fq& a) {
  uint32_t br = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    uint64_t d = (uint64_t)KZG_C(FQ_HALF)[j] - a.v[j] - br;
    br = (uint32_t)(d >> 63);
  }
  return br != 0;
}
KZG_HD bool fq_canon_lt_p(const Fq& a) {
  uint32_t br = 0;
#pragma unroll
  for (int j = 0; j < 12; j++) {
    uint64_t d = (uint64_t)a.v[j] - KZG_C(FQ_P)[j] - br;
    br = (uint32_t)(d >> 63);
  }
  return br != 0;
}

// a^e for a fixed public exponent given as limbs (Fermat inversion, sqrt). Left-to-right,
// 4-bit fixed window.  Montgomery in, Montgomery out.
template <class F, int NE>
KZG_HD void pow_limbs(F& o, const F& a, const uint32_t* e) {
  F tab[16];
  set_one(tab[0]);
  tab[1] = a;
  for (int i = 2; i < 16; i++) mul(tab[i], tab[i - 1], a);
  F acc; set_one(acc);
  bool started = false;
  for (int w = NE * 8 - 1; w >= 0; w--) {
    uint32_t nib = (e[w >> 3] >> ((w & 7) * 4)) & 15;
    if (started) {
      sqr(acc, acc); sqr(acc, acc); sqr(acc, acc); sqr(acc, acc);
    }
    if (nib) {
      if (started) mul(acc, acc, tab[nib]); else { acc = tab[nib]; started = true; }
    }
  }
  o = acc;
}
KZG_HD void inv(Fq& o, const Fq& a) { pow_limbs<Fq, 12>(o, a, KZG_C(FQ_PM2)); }
KZG_HD void inv(Fr& o, const Fr& a) { pow_limbs<Fr, 8>(o, a, KZG_C(FR_PM2)); }

// Inversion by the binary extended Euclid algorithm (shifts and subtractions only): ~25x fewer
// instructions than Fermat's a^(q-2) for a single element, at the price of data-dependent control flow —
// used where ONE thread inverts ONE value (the final to-affine of a commitment).  Montgomery in/out.
KZG_HD void inv_gcd(Fq& o, const Fq& a) {
  if (is_zero(a)) { set_zero(o); return; }
  uint32_t u[12], v[12], x1[12], x2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { u[i] = a.v[i]; v[i] = KZG_C(FQ_P)[i]; x1[i] = 0; x2[i] = 0; }
  x1[0] = 1;
  auto is_one = [](const uint32_t* w) {
    uint32_t acc = w[0] ^ 1u;
    for (int i = 1; i < 12; i++) acc |= w[i];
    return acc == 0;
  };
  auto shr1 = [](uint32_t* w) {
    for (int i = 0; i < 11; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31);
    w[11] >>= 1;
  };
  auto half_mod = [&](uint32_t* x) {          // x <- x/2 mod p ; x + p < 2^382 fits 12 limbs
    if (x[0] & 1u) {
      uint64_t c = 0;
      for (int i = 0; i < 12; i++) { c += (uint64_t)x[i] + KZG_C(FQ_P)[i]; x[i] = (uint32_t)c; c >>= 32; }
    }
    shr1(x);
  };
  auto ge = [](const uint32_t* x, const uint32_t* y) {
    for (int i = 11; i >= 0; i--) { if (x[i] != y[i]) return x[i] > y[i]; }
    return true;
  };
  auto sub_in = [](uint32_t* x, const uint32_t* y) {   // x -= y, x >= y
    uint64_t br = 0;
    for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)x[i] - y[i] - br; x[i] = (uint32_t)d; br = (d >> 63) & 1; }
  };
  auto sub_modp = [&](uint32_t* x, const uint32_t* y) {  // x <- x - y mod p
    uint64_t br = 0;
    for (int i = 0; i < 12; i++) { uint64_t d = (uint64_t)x[i] - y[i] - br; x[i] = (uint32_t)d; br = (d >> 63) & 1; }
    if (br) {
      uint64_t c = 0;
      for (int i = 0; i < 12; i++) { c += (uint64_t)x[i] + KZG_C(FQ_P)[i]; x[i] = (uint32_t)c; c >>= 32; }
    }
  };
  while (!is_one(u) && !is_one(v)) {
    while (!(u[0] & 1u)) { shr1(u); half_mod(x1); }
    while (!(v[0] & 1u)) { shr1(v); half_mod(x2); }
    if (ge(u, v)) { sub_in(u, v); sub_modp(x1, x2); }
    else { sub_in(v, u); sub_modp(x2, x1); }
  }
  Fq z, r2;
  const uint32_t* res = is_one(u) ? x1 : x2;
#pragma unroll
  for (int i = 0; i < 12; i++) { z.v[i] = res[i]; r2.v[i] = KZG_C(FQ_R2)[i]; }
  mul(z, z, r2);   // (aR)^-1 * R^2 * R^-1 = a^-1
  mul(o, z, r2);   // a^-1 * R
}

// x^n for a runtime 64-bit exponent (Fr), plain square-and-multiply; the value equals the
// reference's bespoke Scalar::pow (src/scalar.rs:122-187) because Fr results are canonical.
KZG_HD void pow_u64(Fr& o, const Fr& a, uint64_t n) {
  Fr acc; set_one(acc);
  Fr base = a;
  while (n) {
    if (n & 1) mul(acc, acc, base);
    n >>= 1;
    if (n) sqr(base, base);
  }
  o = acc;
}

// ---- byte I/O ------------------------------------------------------------------------
KZG_HD void fr_load_le(Fr& o, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 8; i++)
    o.v[i] = (uint32_t)b[4 * i] | ((uint32_t)b[4 * i + 1] << 8) | ((uint32_t)b[4 * i + 2] << 16) | ((uint32_t)b[4 * i + 3] << 24);
}
KZG_HD void fr_store_le(uint8_t* b, const Fr& a) {
#pragma unroll
  for (int i = 0; i < 8; i++) {
    b[4 * i] = (uint8_t)a.v[i]; b[4 * i + 1] = (uint8_t)(a.v[i] >> 8);
    b[4 * i + 2] = (uint8_t)(a.v[i] >> 16); b[4 * i + 3] = (uint8_t)(a.v[i] >> 24);
  }
}

}  // namespace kzg
