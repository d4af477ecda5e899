// kzg_b200.hpp — header-only C++ host mirror of the reference's public Rust API over the C ABI.
//
// The reference is compiled code (Rust); with no Rust toolchain in this image the host side above the
// C ABI is C++ with the SAME module / type / method names and error messages:
//   scalar::Scalar                    src/scalar.rs
//   curves::{G1Point, G2Point, bilinear_maps_equal}   src/curves.rs
//   trusted_setup::{SetupArtifactsGenerator, SetupArtifact, TrustedSetup}   src/trusted_setup.rs
//   polynomial::{Polynomial, Evaluation}                                    src/polynomial.rs
// anyhow::Error becomes kzg::Error (what() carries the reference's message).  One FFI call per bulk op.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>
#include "kzg_b200.h"

namespace kzg {

struct Error : std::runtime_error {
  int32_t status;
  Error(int32_t st, const std::string& msg) : std::runtime_error(msg), status(st) {}
};

class Context {
 public:
  explicit Context(int device = 0) {
    int32_t st = kzg_ctx_create(&h_, device);
    if (st != KZG_OK) throw Error(st, "kzg_ctx_create failed: no usable CUDA device (there is no CPU fallback)");
  }
  // several devices driven from this process (kzg_ctx_create_multi): the same calls shard internally
  explicit Context(const std::vector<int32_t>& devices) {
    int32_t st = kzg_ctx_create_multi(&h_, devices.data(), (int32_t)devices.size());
    if (st != KZG_OK) throw Error(st, "kzg_ctx_create_multi failed: no usable CUDA devices or no peer access between them");
  }
  int device_count() const { return kzg_ctx_device_count(h_); }
  ~Context() { kzg_ctx_destroy(h_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  kzg_ctx* get() const { return h_; }
  void check(int32_t st) const {
    if (st != KZG_OK) {
      std::string m = kzg_last_error(h_);
      throw Error(st, m.empty() ? kzg_status_string(st) : m);
    }
  }
 private:
  kzg_ctx* h_ = nullptr;
};

namespace scalar {
class Scalar {
 public:
  Scalar() : le_{} {}
  static Scalar from_le_bytes(const std::array<uint8_t, 32>& b) { Scalar s; kzg_fr_reduce(b.data(), s.le_.data()); return s; }
  static Scalar from_be_bytes(const std::array<uint8_t, 32>& b) {
    std::array<uint8_t, 32> le; for (int i = 0; i < 32; i++) le[i] = b[31 - i];
    return from_le_bytes(le);
  }
  // impl From<i128>: a > 0 -> a, a <= 0 -> r - |a| (src/scalar.rs:27-48)
  static Scalar from_i128(__int128 a) {
    unsigned __int128 mag = a < 0 ? (unsigned __int128)(-(a + 1)) + 1 : (unsigned __int128)a;
    std::array<uint8_t, 32> le{}; for (int i = 0; i < 16; i++) le[i] = (uint8_t)(mag >> (8 * i));
    Scalar s = from_le_bytes(le);
    return a > 0 ? s : s.neg();
  }
  std::array<uint8_t, 32> to_le_bytes() const { return le_; }
  std::array<uint8_t, 32> to_be_bytes() const { std::array<uint8_t, 32> b; for (int i = 0; i < 32; i++) b[i] = le_[31 - i]; return b; }
  Scalar mul(const Scalar& o) const { Scalar r; kzg_fr_mul(le_.data(), o.le_.data(), r.le_.data()); return r; }
  Scalar add(const Scalar& o) const { Scalar r; kzg_fr_add(le_.data(), o.le_.data(), r.le_.data()); return r; }
  Scalar sub(const Scalar& o) const { Scalar r; kzg_fr_sub(le_.data(), o.le_.data(), r.le_.data()); return r; }
  Scalar neg() const { Scalar r; kzg_fr_neg(le_.data(), r.le_.data()); return r; }
  Scalar pow(size_t n) const { Scalar r; kzg_fr_pow(le_.data(), n, r.le_.data()); return r; }
  bool is_zero() const { for (auto b : le_) if (b) return false; return true; }
  bool operator==(const Scalar& o) const { return le_ == o.le_; }
  bool operator!=(const Scalar& o) const { return !(*this == o); }
  // impl Display: base-10 string of the canonical value (src/scalar.rs:272-345)
  std::string to_string() const {
    std::array<uint8_t, 32> q = le_;
    std::string digits;
    bool nonzero = true;
    while (nonzero) {
      unsigned rem = 0; nonzero = false;
      for (int i = 31; i >= 0; i--) { unsigned cur = (rem << 8) | q[i]; q[i] = (uint8_t)(cur / 10); rem = cur % 10; if (q[i]) nonzero = true; }
      digits.push_back((char)('0' + rem));
    }
    return std::string(digits.rbegin(), digits.rend());
  }
 private:
  std::array<uint8_t, 32> le_;
};
}  // namespace scalar

namespace curves {
class G1Point {
 public:
  // Deserialize (src/curves.rs:112-183)
  static G1Point from_compressed(const uint8_t* b, size_t len) {
    if (len != 48) throw Error(KZG_E_BAD_LENGTH, "Invalid length, expected 48, got " + std::to_string(len));
    if (kzg_g1_check(b) != KZG_OK) throw Error(KZG_E_BAD_POINT, "Got error while uncompressing");
    G1Point p; std::copy(b, b + 48, p.c_.begin()); return p;
  }
  static G1Point from_scalar(const scalar::Scalar& a) { G1Point p; kzg_g1_from_scalar(a.to_le_bytes().data(), p.c_.data()); return p; }
  static G1Point from_i128(__int128 a) {
    G1Point p = from_scalar(scalar::Scalar::from_i128(a < 0 ? -a : a));
    return a < 0 ? from_scalar(scalar::Scalar()).sub(p) : p;
  }
  G1Point add(const G1Point& b) const { G1Point r; kzg_g1_add(c_.data(), b.c_.data(), r.c_.data()); return r; }
  G1Point sub(const G1Point& b) const { G1Point r; kzg_g1_sub(c_.data(), b.c_.data(), r.c_.data()); return r; }
  G1Point mult(const scalar::Scalar& a) const { G1Point r; kzg_g1_mul(c_.data(), a.to_le_bytes().data(), r.c_.data()); return r; }
  const std::array<uint8_t, 48>& compressed() const { return c_; }   // Serialize (src/curves.rs:99-110)
  bool operator==(const G1Point& o) const { return c_ == o.c_; }
  static G1Point raw(const uint8_t* b) { G1Point p; std::copy(b, b + 48, p.c_.begin()); return p; }
 private:
  std::array<uint8_t, 48> c_{};
};

// src/curves.rs:185-343
class G2Point {
 public:
  static G2Point from_compressed(const uint8_t* b, size_t len) {
    if (len != 96) throw Error(KZG_E_BAD_LENGTH, "Invalid length, expected 96, got " + std::to_string(len));
    if (kzg_g2_check(b) != KZG_OK) throw Error(KZG_E_BAD_POINT, "Got error while uncompressing");
    G2Point p; std::copy(b, b + 96, p.c_.begin()); return p;
  }
  static G2Point from_scalar(const scalar::Scalar& a) { G2Point p; kzg_g2_from_scalar(a.to_le_bytes().data(), p.c_.data()); return p; }
  static G2Point from_i128(__int128 a) {
    G2Point p = from_scalar(scalar::Scalar::from_i128(a < 0 ? -a : a));
    if (a < 0) { G2Point n; kzg_g2_neg(p.c_.data(), n.c_.data()); return n; }
    return p;
  }
  G2Point sub(const G2Point& b) const { G2Point r; kzg_g2_sub(c_.data(), b.c_.data(), r.c_.data()); return r; }
  const std::array<uint8_t, 96>& compressed() const { return c_; }   // Serialize (src/curves.rs:260-271)
  bool operator==(const G2Point& o) const { return c_ == o.c_; }
  static G2Point raw(const uint8_t* b) { G2Point p; std::copy(b, b + 96, p.c_.begin()); return p; }
 private:
  std::array<uint8_t, 96> c_{};
};

// bilinear_map(p1, q1) == bilinear_map(p2, q2) (src/curves.rs:355-371; the only use of the pairing,
// src/polynomial.rs:293)
inline bool bilinear_maps_equal(const G1Point& p1, const G2Point& q1, const G1Point& p2, const G2Point& q2) {
  int32_t eq = 0;
  int32_t st = kzg_pairing_equal(p1.compressed().data(), q1.compressed().data(), p2.compressed().data(), q2.compressed().data(), &eq);
  if (st != KZG_OK) throw Error(st, kzg_status_string(st));
  return eq != 0;
}
}  // namespace curves

namespace trusted_setup {
struct SetupArtifact { curves::G1Point g1; curves::G2Point g2; bool has_g2 = false; };

// Vec<SetupArtifact> + its device-resident copy
class TrustedSetup {
 public:
  TrustedSetup(std::shared_ptr<Context> ctx, kzg_setup* h) : ctx_(std::move(ctx)), h_(h) {}
  ~TrustedSetup() { kzg_setup_destroy(h_); }
  TrustedSetup(const TrustedSetup&) = delete;
  size_t len() const { return kzg_setup_len(h_); }
  size_t g2_len() const { return kzg_setup_g2_len(h_); }
  SetupArtifact at(size_t i) const {
    if (i >= len()) throw std::out_of_range("index out of bounds: the len is " + std::to_string(len()) + " but the index is " + std::to_string(i));
    uint8_t b[48]; ctx_->check(kzg_setup_export_g1(ctx_->get(), h_, i, 1, b));
    SetupArtifact a; a.g1 = curves::G1Point::raw(b);
    if (i < g2_len()) { uint8_t c[96]; ctx_->check(kzg_setup_export_g2(ctx_->get(), h_, i, 1, c)); a.g2 = curves::G2Point::raw(c); a.has_g2 = true; }
    return a;
  }
  // Vec<SetupArtifact> from deserialised artifacts (the CLI's setup.json path)
  static std::unique_ptr<TrustedSetup> from_artifacts(std::shared_ptr<Context> ctx, const std::vector<SetupArtifact>& items) {
    std::vector<uint8_t> g1(items.size() * 48), g2;
    for (size_t i = 0; i < items.size(); i++) std::copy(items[i].g1.compressed().begin(), items[i].g1.compressed().end(), g1.begin() + 48 * i);
    for (size_t i = 0; i < items.size() && items[i].has_g2; i++) g2.insert(g2.end(), items[i].g2.compressed().begin(), items[i].g2.compressed().end());
    kzg_setup* h = nullptr;
    ctx->check(kzg_setup_import_g1(ctx->get(), g1.data(), items.size(), &h));
    auto out = std::make_unique<TrustedSetup>(ctx, h);
    if (!g2.empty()) ctx->check(kzg_setup_import_g2(ctx->get(), h, g2.data(), g2.size() / 96));
    return out;
  }
  void prepare_msm(int window_bits = 0) { ctx_->check(kzg_setup_prepare_msm(ctx_->get(), h_, window_bits)); }
  void prepare_fixed(size_t n_points, int window_bits = 0) { ctx_->check(kzg_setup_prepare_fixed(ctx_->get(), h_, n_points, window_bits)); }
  const std::shared_ptr<Context>& ctx() const { return ctx_; }
  kzg_setup* handle() const { return h_; }
 private:
  std::shared_ptr<Context> ctx_;
  kzg_setup* h_;
};

class SetupArtifactsGenerator {
 public:
  SetupArtifactsGenerator(std::shared_ptr<Context> ctx, const std::array<uint8_t, 32>& secret_be) : ctx_(std::move(ctx)), secret_(secret_be) {}
  // .take(n).collect()  (src/trusted_setup.rs:40-78)
  // g2_count: how many items also get their G2 point (the reference: all of them, src/trusted_setup.rs:64-72)
  std::unique_ptr<TrustedSetup> take(size_t n, size_t g2_count = SIZE_MAX) const {
    kzg_setup* h = nullptr;
    ctx_->check(kzg_setup_generate(ctx_->get(), secret_.data(), 0, n, &h));
    auto out = std::make_unique<TrustedSetup>(ctx_, h);
    ctx_->check(kzg_setup_generate_g2(ctx_->get(), h, secret_.data(), g2_count < n ? g2_count : n));
    return out;
  }
 private:
  std::shared_ptr<Context> ctx_;
  std::array<uint8_t, 32> secret_;
};
}  // namespace trusted_setup

namespace polynomial {
using scalar::Scalar;
class Evaluation;

class Polynomial {
 public:
  // TryFrom<Vec<Scalar>>: trailing zeros trimmed, an all-zero input keeps one coefficient (src/polynomial.rs:55-75)
  static Polynomial try_from(std::vector<Scalar> v) {
    size_t last = 0;
    for (size_t i = 0; i < v.size(); i++) if (!v[i].is_zero()) last = i;
    if (!v.empty()) v.resize(last + 1);
    return Polynomial(std::move(v));
  }
  static Polynomial try_from_i128(const std::vector<__int128>& v) {
    std::vector<Scalar> s; for (auto a : v) s.push_back(Scalar::from_i128(a));
    return try_from(std::move(s));
  }
  static Polynomial from(const Scalar& c) { return Polynomial(c.is_zero() ? std::vector<Scalar>{} : std::vector<Scalar>{c}); }
  // derived Deserialize keeps the coefficients as stored: no re-trimming (src/polynomial.rs:9-12)
  static Polynomial from_raw_coefficients(std::vector<Scalar> v) { return Polynomial(std::move(v)); }
  // Polynomial::sub (src/polynomial.rs:128-145)
  Polynomial sub(const Polynomial& o) const {
    std::vector<Scalar> co;
    if (coefficients_.size() > o.coefficients_.size()) {
      co = coefficients_;
      for (size_t i = 0; i < o.coefficients_.size(); i++) co[i] = co[i].sub(o.coefficients_[i]);
    } else {
      for (auto& c : o.coefficients_) co.push_back(c.neg());
      for (size_t i = 0; i < coefficients_.size(); i++) co[i] = coefficients_[i].add(co[i]);
    }
    return try_from(std::move(co));
  }
  // Polynomial::divide_by_root (src/polynomial.rs:150-195): one kzg_quotient call
  Polynomial divide_by_root(const Context& ctx, const Scalar& root) const {
    if (coefficients_.empty()) return Polynomial({});
    if (coefficients_.size() == 1) {
      if (coefficients_[0].is_zero()) return Polynomial({});
      throw Error(KZG_E_CONSTANT_POLY, kzg_status_string(KZG_E_CONSTANT_POLY));
    }
    auto b = bytes(); std::vector<uint8_t> q(coefficients_.size() * 32); uint64_t ql = 0; std::array<uint8_t, 32> zero{};
    ctx.check(kzg_quotient(ctx.get(), b.data(), coefficients_.size(), root.to_le_bytes().data(), zero.data(), q.data(), &ql));
    std::vector<Scalar> co;
    for (uint64_t i = 0; i < ql; i++) { std::array<uint8_t, 32> le; std::copy(q.begin() + 32 * i, q.begin() + 32 * i + 32, le.begin()); co.push_back(Scalar::from_le_bytes(le)); }
    return try_from(std::move(co));
  }
  // impl Display (src/polynomial.rs:218-247)
  std::string to_string() const {
    if (coefficients_.empty()) return "0";
    auto term = [](const Scalar& c, size_t d) { return c.to_string() + (d == 0 ? "" : d == 1 ? "x" : "x^" + std::to_string(d)); };
    size_t n = coefficients_.size();
    std::string out = term(coefficients_[n - 1], n - 1);
    for (size_t i = n - 1; i-- > 0;) if (!coefficients_[i].is_zero()) out += " + " + term(coefficients_[i], i);
    return out;
  }
  uint32_t degree() const { return coefficients_.empty() ? 0 : (uint32_t)(coefficients_.size() - 1); }
  const std::vector<Scalar>& coefficients() const { return coefficients_; }

  inline Evaluation evaluate(const Context& ctx, const Scalar& x) const;
  // Polynomial::commit (src/polynomial.rs:200-215)
  curves::G1Point commit(const trusted_setup::TrustedSetup& setup) const {
    auto b = bytes(); uint8_t out[48];
    setup.ctx()->check(kzg_commit(setup.ctx()->get(), setup.handle(), b.data(), coefficients_.size(), out));
    return curves::G1Point::raw(out);
  }
  std::vector<uint8_t> bytes() const {
    std::vector<uint8_t> b(coefficients_.size() * 32);
    for (size_t i = 0; i < coefficients_.size(); i++) { auto le = coefficients_[i].to_le_bytes(); std::copy(le.begin(), le.end(), b.begin() + 32 * i); }
    return b;
  }
 private:
  explicit Polynomial(std::vector<Scalar> c) : coefficients_(std::move(c)) {}
  std::vector<Scalar> coefficients_;
};

class Evaluation {
 public:
  Scalar point, result;
  // Evaluation::generate_proof (src/polynomial.rs:260-269)
  curves::G1Point generate_proof(const Polynomial& p, const trusted_setup::TrustedSetup& setup) const {
    auto b = p.bytes(); uint8_t y[32], out[48];
    setup.ctx()->check(kzg_open(setup.ctx()->get(), setup.handle(), b.data(), p.coefficients().size(), point.to_le_bytes().data(),
                                result.to_le_bytes().data(), y, out));
    return curves::G1Point::raw(out);
  }
  // Evaluation::verify_proof (src/polynomial.rs:276-294); setup[1] must exist (the reference indexes it unguarded)
  bool verify_proof(const curves::G1Point& proof, const curves::G1Point& commitment, const trusted_setup::TrustedSetup& setup) const {
    trusted_setup::SetupArtifact a1 = setup.at(1);
    if (!a1.has_g2) throw Error(KZG_E_SETUP_TOO_SHORT, "setup item 1 carries no G2 point");
    int32_t ok = 0;
    int32_t st = kzg_verify_proof(proof.compressed().data(), commitment.compressed().data(), point.to_le_bytes().data(),
                                  result.to_le_bytes().data(), a1.g2.compressed().data(), &ok);
    if (st != KZG_OK) throw Error(st, kzg_status_string(st));
    return ok != 0;
  }
};

inline Evaluation Polynomial::evaluate(const Context& ctx, const Scalar& x) const {
  auto b = bytes(); std::array<uint8_t, 32> y;
  ctx.check(kzg_evaluate(ctx.get(), b.data(), coefficients_.size(), x.to_le_bytes().data(), y.data()));
  return Evaluation{x, Scalar::from_le_bytes(y)};
}
}  // namespace polynomial

}  // namespace kzg
