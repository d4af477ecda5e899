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

// Halve the range of a canonical scalar before the signed-digit recoding: for k > (r-1)/2 continue with r - k and
// flip the sign of every digit ([k]P = -[r-k]P).  Afterwards k < 2^254, so ceil(255/c) windows of c bits hold it
// together with the top carry (windows_for, recode.cuh) - one window fewer when c divides 255 (c = 15, 17).
// Returns true when the scalar was replaced by r - k.
KZG_HD bool fr_fold_half(Fr& k) {
  uint32_t t[8]; uint32_t br = 0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    uint64_t d = (uint64_t)KZG_C(FR_P)[j] - k.v[j] - br;
    t[j] = (uint32_t)d; br = (uint32_t)(d >> 63);
  }
  br = 0;                       // r - k < k ?
#pragma unroll
  for (int j = 0; j < 8; j++) {
    uint64_t d = (uint64_t)t[j] - k.v[j] - br;
    br = (uint32_t)(d >> 63);
  }
  if (!br) return false;
#pragma unroll
  for (int j = 0; j < 8; j++) k.v[j] = t[j];
  return true;
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

// Inversion by the optimised binary GCD of Pornin ("Optimized Binary GCD for Modular Inversion", 2020):
// the 381-bit operands a, b are replaced, 30 steps at a time, by 62-bit approximations (their top 32 bits
// and their exact low 30 bits); the 30 plain binary-GCD steps run on those two words and yield a 2x2
// integer matrix (|f|+|g| <= 2^30) that is then applied once to (a, b) — exact division by 2^30 — and to
// the Bezout pair (u, v) — division by 2^30 modulo q, Montgomery style.  28 outer iterations cover the
// 2*381 - 1 steps a 381-bit modulus can need; control flow is identical for every input, so a warp of
// independent inversions never diverges.  ~12x fewer instructions than Fermat's a^(q-2): used for the
// final to-affine of a commitment, the top of the grid-wide batch inversions (batch_tree.cu), Fq2
// inversion.  Montgomery in, Montgomery out; inv_gcd(0) = 0.
namespace bingcd {
// t = +-fa * a +- ga * b as a 13-limb two's-complement integer (fa, ga < 2^31 magnitudes, signs sf, sg)
KZG_HD void lin13(uint32_t* t, uint32_t fa, bool sf, const uint32_t* a, uint32_t ga, bool sg, const uint32_t* b) {
  uint32_t pa[13], pb[13];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) { c += (uint64_t)fa * a[i]; pa[i] = (uint32_t)c; c >>= 32; }
  pa[12] = (uint32_t)c;
  c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) { c += (uint64_t)ga * b[i]; pb[i] = (uint32_t)c; c >>= 32; }
  pb[12] = (uint32_t)c;
  const uint32_t ma = sf ? 0xFFFFFFFFu : 0u, mb = sg ? 0xFFFFFFFFu : 0u;
  uint64_t ca = sf ? 1u : 0u, cb = sg ? 1u : 0u, cs = 0;
#pragma unroll
  for (int i = 0; i < 13; i++) {
    ca += (uint64_t)(pa[i] ^ ma); uint32_t xa = (uint32_t)ca; ca >>= 32;      // conditional negation
    cb += (uint64_t)(pb[i] ^ mb); uint32_t xb = (uint32_t)cb; cb >>= 32;
    cs += (uint64_t)xa + xb; t[i] = (uint32_t)cs; cs >>= 32;
  }
}
KZG_HD void neg13(uint32_t* t) {
  uint64_t c = 1;
#pragma unroll
  for (int i = 0; i < 13; i++) { c += (uint64_t)(~t[i]); t[i] = (uint32_t)c; c >>= 32; }
}
// bits [sh, sh+32) of a 384-bit integer, sh in [30, 352], without dynamic register indexing
KZG_HD uint32_t extract32(const uint32_t* a, uint32_t sh) {
  const uint32_t w = sh >> 5, off = sh & 31;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    lo = (i == (int)w) ? a[i] : lo;
    hi = (i == (int)w + 1) ? a[i] : hi;
  }
  return off ? ((lo >> off) | (hi << (32 - off))) : lo;
}
// dst = t / 2^30 mod q for a 13-limb two's-complement t in (-2^30 q, 2^30 q): add the multiple of q that
// clears the low 30 bits (Montgomery step), shift, fold the result from (-q, 2q) into [0, q)
KZG_HD void div30_mod_q(uint32_t* dst, uint32_t* t) {
  const uint32_t m = (t[0] * FQ_NEGINV30) & 0x3FFFFFFFu;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) { c += (uint64_t)m * KZG_C(FQ_P)[i] + t[i]; t[i] = (uint32_t)c; c >>= 32; }
  t[12] = (uint32_t)(c + t[12]);
  const bool negative = (t[12] >> 31) != 0;
  uint32_t r[12], w[12], s2[12];
#pragma unroll
  for (int i = 0; i < 12; i++) r[i] = (t[i] >> 30) | (t[i + 1] << 2);
  uint64_t cc = 0; uint32_t br = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    cc += (uint64_t)r[i] + KZG_C(FQ_P)[i]; w[i] = (uint32_t)cc; cc >>= 32;
    uint64_t d = (uint64_t)r[i] - KZG_C(FQ_P)[i] - br; s2[i] = (uint32_t)d; br = (uint32_t)(d >> 63);
  }
  const bool ge = !negative && br == 0;
#pragma unroll
  for (int i = 0; i < 12; i++) dst[i] = negative ? w[i] : (ge ? s2[i] : r[i]);
}
}  // namespace bingcd

KZG_HD void inv_gcd(Fq& o, const Fq& x) {
  uint32_t a[12], b[12], u[12], v[12];
#pragma unroll
  for (int i = 0; i < 12; i++) { a[i] = x.v[i]; b[i] = KZG_C(FQ_P)[i]; u[i] = 0; v[i] = 0; }
  u[0] = 1;
#pragma unroll 1
  for (int it = 0; it < 28; it++) {
    // ---- 62-bit approximations: n = max(len(a), len(b), 62); top 32 bits above bit n-32, low 30 bits exact
    uint32_t top = 0, topw = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
      uint32_t c = a[i] | b[i];
      if (c) { top = c; topw = (uint32_t)i; }
    }
    uint32_t blen = 0;
    for (uint32_t t = top; t; t >>= 1) blen++;               // <= 32 steps; host + device
    uint32_t n = topw * 32 + blen;
    if (n < 62) n = 62;
    const uint32_t sh = n - 32;
    uint64_t xa = ((uint64_t)bingcd::extract32(a, sh) << 30) | (a[0] & 0x3FFFFFFFu);
    uint64_t xb = ((uint64_t)bingcd::extract32(b, sh) << 30) | (b[0] & 0x3FFFFFFFu);
    // ---- 30 binary-GCD steps on the approximations
    int32_t f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll 1
    for (int i = 0; i < 30; i++) {
      const bool odd = (xa & 1u) != 0;
      const bool swp = odd && (xa < xb);
      if (swp) {
        uint64_t tx = xa; xa = xb; xb = tx;
        int32_t tf = f0; f0 = f1; f1 = tf;
        int32_t tg = g0; g0 = g1; g1 = tg;
      }
      if (odd) { xa -= xb; f0 -= f1; g0 -= g1; }
      xa >>= 1;
      f1 <<= 1; g1 <<= 1;
    }
    // ---- (a, b) <- (f0 a + g0 b, f1 a + g1 b) / 2^30, made non-negative
    uint32_t ta[13], tb[13];
    bingcd::lin13(ta, (uint32_t)(f0 < 0 ? -f0 : f0), f0 < 0, a, (uint32_t)(g0 < 0 ? -g0 : g0), g0 < 0, b);
    bingcd::lin13(tb, (uint32_t)(f1 < 0 ? -f1 : f1), f1 < 0, a, (uint32_t)(g1 < 0 ? -g1 : g1), g1 < 0, b);
    if (ta[12] >> 31) { bingcd::neg13(ta); f0 = -f0; g0 = -g0; }
    if (tb[12] >> 31) { bingcd::neg13(tb); f1 = -f1; g1 = -g1; }
#pragma unroll
    for (int i = 0; i < 12; i++) {
      a[i] = (ta[i] >> 30) | (ta[i + 1] << 2);
      b[i] = (tb[i] >> 30) | (tb[i + 1] << 2);
    }
    // ---- (u, v) <- the same combination / 2^30 mod q
    uint32_t tu[13], tv[13];
    bingcd::lin13(tu, (uint32_t)(f0 < 0 ? -f0 : f0), f0 < 0, u, (uint32_t)(g0 < 0 ? -g0 : g0), g0 < 0, v);
    bingcd::lin13(tv, (uint32_t)(f1 < 0 ? -f1 : f1), f1 < 0, u, (uint32_t)(g1 < 0 ? -g1 : g1), g1 < 0, v);
    bingcd::div30_mod_q(u, tu);
    bingcd::div30_mod_q(v, tv);
  }
  // gcd reached: b = 1, v = x^-1 (as integers mod q); x = a_mont = aR  ->  v = a^-1 R^-1
  Fq z, r2;
#pragma unroll
  for (int i = 0; i < 12; i++) { z.v[i] = v[i]; r2.v[i] = KZG_C(FQ_R2)[i]; }
  if (is_zero(x)) { set_zero(o); return; }
  mul(z, z, r2);   // a^-1
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
