// C++ host-API smoke test: the reference's degree-1 bench case through include/kzg_b200.hpp.
// Build: g++ -std=c++17 -Iinclude tests/cpp/test_api.cpp -o tests/cpp/test_api.bin kzg_poly_commit_exploration_b200/libkzg_b200.so
// CPU part (no GPU): Scalar / G1Point helpers.  GPU part runs when a device exists.
#include <cstdio>
#include <cstring>
#include "kzg_b200.hpp"
using namespace kzg;
static std::string hex(const uint8_t* b, size_t n) { static const char* d = "0123456789abcdef"; std::string s; for (size_t i = 0; i < n; i++) { s += d[b[i] >> 4]; s += d[b[i] & 15]; } return s; }
int main() {
  using scalar::Scalar; using curves::G1Point;
  int fails = 0;
  auto expect = [&](bool ok, const char* what) { if (!ok) { printf("FAIL %s\n", what); fails++; } };
  Scalar five = Scalar::from_i128(5), ten = Scalar::from_i128(10);
  expect(five.pow(2).add(ten) == Scalar::from_i128(35), "pow/add");
  expect(Scalar::from_i128(-1).add(Scalar::from_i128(1)).is_zero(), "neg");
  G1Point g = G1Point::from_i128(1);
  expect(hex(g.compressed().data(), 48) == "97f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb", "generator");
  expect(g.add(g) == G1Point::from_i128(2), "G+G");
  expect(hex(G1Point::from_i128(-1).compressed().data(), 48).substr(0, 2) == "b7", "-G");
  std::shared_ptr<Context> ctx;
  try { ctx = std::make_shared<Context>(0); } catch (const Error& e) { printf("no GPU (%s): CPU part %s\n", e.what(), fails ? "FAILED" : "ok"); return fails; }
  std::array<uint8_t, 32> secret; for (int i = 0; i < 32; i++) secret[i] = (uint8_t)i;   // benches/polynomial_commitment.rs:17-19
  auto setup = trusted_setup::SetupArtifactsGenerator(ctx, secret).take(2);
  auto poly = polynomial::Polynomial::try_from({five.pow(0).add(ten), five.pow(1).add(ten)});   // c_i = 5^i + 10
  Scalar z = five.pow(1).add(Scalar::from_i128(20));
  auto c = poly.commit(*setup);
  auto ev = poly.evaluate(*ctx, z);
  auto proof = ev.generate_proof(poly, *setup);
  expect(hex(c.compressed().data(), 48) == "add892451b1c924699eaadfa1566680ac35623081bd4c71db9aa61019f80283e7ab2f173a9d68b11d9defeb998b988bd", "commit KAT");
  expect(hex(ev.result.to_le_bytes().data(), 32) == "8201000000000000000000000000000000000000000000000000000000000000", "y KAT");
  expect(hex(proof.compressed().data(), 48) == "8d9e19b3f4c7c233a6112e5397309f9812a4f61f754f11dd3dcb8b07d55a7b1dfea65f19a1488a14fef9a41495083582", "proof KAT");
  try { polynomial::Evaluation{z, ev.result.add(five)}.generate_proof(poly, *setup); expect(false, "wrong result accepted"); }
  catch (const Error& e) { expect(e.status == KZG_E_REMAINDER_NONZERO && std::string(e.what()).find("constant terms do not add up") != std::string::npos, "error message"); }
  printf("C++ API test %s\n", fails ? "FAILED" : "ok");
  return fails;
}
