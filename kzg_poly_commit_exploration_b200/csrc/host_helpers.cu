// Single-element host helpers of the C ABI (include/kzg_b200.h): the per-element methods of
// the reference's public API — Scalar::{mul,add,sub,neg,pow} (src/scalar.rs:111-223) and
// G1Point::{add,sub,mult,from_scalar} (src/curves.rs:50-96).  They run the host
// instantiation of the SAME field / curve formulas the kernels use (field.cuh, g1.cuh), which
// is also how the CPU test-suite checks those formulas against the oracle.
#include "../../include/kzg_b200.h"
#include "g1.cuh"

using namespace kzg;

namespace {
void load_fr_mont(Fr& o, const uint8_t* b) { Fr t; fr_load_le(t, b); to_mont(o, t); }
void store_fr_mont(uint8_t* b, const Fr& a) { Fr t; from_mont(t, a); fr_store_le(b, t); }
void compress_xyzz(uint8_t* out, const G1Xyzz& p) { G1Affine a; to_affine(a, p); compress(out, a); }
}  // namespace

extern "C" {

void kzg_fr_mul(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  Fr x, y, z; load_fr_mont(x, a); load_fr_mont(y, b); mul(z, x, y); store_fr_mont(out, z);
}
void kzg_fr_add(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  Fr x, y, z; load_fr_mont(x, a); load_fr_mont(y, b); add(z, x, y); store_fr_mont(out, z);
}
void kzg_fr_sub(const uint8_t a[32], const uint8_t b[32], uint8_t out[32]) {
  Fr x, y, z; load_fr_mont(x, a); load_fr_mont(y, b); sub(z, x, y); store_fr_mont(out, z);
}
void kzg_fr_neg(const uint8_t a[32], uint8_t out[32]) {
  Fr x, z; load_fr_mont(x, a); neg(z, x); store_fr_mont(out, z);
}
void kzg_fr_pow(const uint8_t a[32], uint64_t n, uint8_t out[32]) {
  Fr x, z; load_fr_mont(x, a); pow_u64(z, x, n); store_fr_mont(out, z);
}
void kzg_fr_reduce(const uint8_t a[32], uint8_t out[32]) {
  Fr x; fr_load_le(x, a); fr_canon_reduce(x); fr_store_le(out, x);
}

int32_t kzg_g1_check(const uint8_t p[48]) {
  G1Affine a;
  return uncompress(a, p) ? KZG_E_BAD_POINT : KZG_OK;
}
int32_t kzg_g1_add(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) {
  G1Affine pa, pb;
  if (uncompress(pa, a) || uncompress(pb, b)) return KZG_E_BAD_POINT;
  G1Xyzz acc; from_affine(acc, pa); madd(acc, pb, false);
  compress_xyzz(out, acc);
  return KZG_OK;
}
int32_t kzg_g1_sub(const uint8_t a[48], const uint8_t b[48], uint8_t out[48]) {
  G1Affine pa, pb;
  if (uncompress(pa, a) || uncompress(pb, b)) return KZG_E_BAD_POINT;
  G1Xyzz acc; from_affine(acc, pa); madd(acc, pb, true);
  compress_xyzz(out, acc);
  return KZG_OK;
}
int32_t kzg_g1_mul(const uint8_t p[48], const uint8_t k_le[32], uint8_t out[48]) {
  G1Affine pa;
  if (uncompress(pa, p)) return KZG_E_BAD_POINT;
  // G1Point::mult consumes Scalar::to_le_bytes, i.e. the canonical residue (src/curves.rs:90-96)
  Fr k; fr_load_le(k, k_le); fr_canon_reduce(k);
  G1Xyzz acc; scalar_mul(acc, pa, k.v, 256);
  compress_xyzz(out, acc);
  return KZG_OK;
}
void kzg_g1_from_scalar(const uint8_t k_le[32], uint8_t out[48]) {
  G1Affine g; g1_generator(g);
  Fr k; fr_load_le(k, k_le); fr_canon_reduce(k);
  G1Xyzz acc; scalar_mul(acc, g, k.v, 256);
  compress_xyzz(out, acc);
}

const char* kzg_status_string(int32_t st) {
  switch (st) {
    case KZG_OK: return "ok";
    case KZG_E_DEGREE_TOO_HIGH:
      return "Setup does not allow for commitment generation of the polynomial. The polynomial degree is too high.";
    case KZG_E_CONSTANT_POLY: return "Unable to divide a constant polynomial";
    case KZG_E_REMAINDER_NONZERO:
      return "[divide_by_root] Fail to divide the polynomial by a root, constant terms do not add up";
    case KZG_E_BAD_POINT: return "Got error while uncompressing";
    case KZG_E_BAD_LENGTH: return "Invalid length";
    case KZG_E_SETUP_TOO_SHORT: return "setup too short";
    case KZG_E_BAD_ARG: return "bad argument";
    case KZG_E_CUDA: return "CUDA error";
    case KZG_E_NOMEM: return "out of device memory";
    default: return "unknown status";
  }
}

}  // extern "C"
