// Single-process multi-GPU test of the C ABI (kzg_ctx_create_multi): ONE host process commits n coefficients on all
// visible GPUs and must get, bit for bit, what one GPU returns — the shape of the reference, which is a single process
// (src/main.rs:170, benches/polynomial_commitment.rs:38-47).
//   usage: test_multi.bin [log2 n = 20] [devices = all visible; "0,0" shares one GPU between two members]
// Covers: range-sharded setup generation (exported items equal the 1-GPU items), commit / msm with calls in flight,
// open (quotient on member 0, sum over all ranges), replicated blob batch split by row, error parity.
// Build: g++ -std=c++17 -O2 -Iinclude tests/cpp/test_multi.cpp -o tests/cpp/test_multi.bin -Lkzg_poly_commit_exploration_b200 -l:libkzg_b200.so
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "kzg_b200.h"

static int fails = 0;
#define EXPECT(cond, what) do { if (!(cond)) { printf("FAIL %s (line %d)\n", what, __LINE__); fails++; } } while (0)
#define OK(ctx, call) do { int32_t st__ = (call); if (st__ != KZG_OK) { printf("FAIL %s -> %d: %s (line %d)\n", #call, st__, kzg_last_error(ctx), __LINE__); return 1; } } while (0)

// U(seed, k) of SURVEY 8d: four SplitMix64 words, masked to 255 bits, minus r if >= r
static uint64_t mix(uint64_t z) { z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }
static void scalar_u(uint64_t seed, uint64_t k, uint8_t out[32]) {
  static const uint64_t R[4] = {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull};
  uint64_t w[4];
  for (int j = 0; j < 4; j++) w[j] = mix(seed + (4 * k + (uint64_t)j + 1) * 0x9E3779B97F4A7C15ull);
  w[3] &= 0x7fffffffffffffffull;
  bool ge = true;
  for (int j = 3; j >= 0; j--) { if (w[j] != R[j]) { ge = w[j] > R[j]; break; } }
  if (ge) { unsigned __int128 b = 0; for (int j = 0; j < 4; j++) { unsigned __int128 d = (unsigned __int128)w[j] - R[j] - (uint64_t)b; w[j] = (uint64_t)d; b = (d >> 64) & 1; } }
  memcpy(out, w, 32);
}
static std::string hex(const uint8_t* b, size_t n) { static const char* d = "0123456789abcdef"; std::string s; for (size_t i = 0; i < n; i++) { s += d[b[i] >> 4]; s += d[b[i] & 15]; } return s; }

int main(int argc, char** argv) {
  const int logn = argc > 1 ? atoi(argv[1]) : 20;
  const uint64_t n = 1ull << logn;
  std::vector<int32_t> devs;
  if (argc > 2) { for (char* t = strtok(argv[2], ","); t; t = strtok(nullptr, ",")) devs.push_back(atoi(t)); }
  kzg_ctx* one = nullptr;
  if (kzg_ctx_create(&one, 0) != KZG_OK) { printf("no GPU: nothing to test\n"); return 0; }
  if (devs.empty()) {
    // count the devices by probing (the C ABI deliberately has no CUDA types in it)
    for (int d = 0; d < 32; d++) { kzg_ctx* c = nullptr; if (kzg_ctx_create(&c, d) != KZG_OK) break; kzg_ctx_destroy(c); devs.push_back(d); }
  }
  kzg_ctx* grp = nullptr;
  OK(one, kzg_ctx_create_multi(&grp, devs.data(), (int32_t)devs.size()));
  printf("group of %d device(s), n = 2^%d\n", kzg_ctx_device_count(grp), logn);

  uint8_t secret[32]; for (int i = 0; i < 32; i++) secret[i] = (uint8_t)i;       // benches/polynomial_commitment.rs:17-19
  // page-locked scalars (kzg_host_alloc): the copies of the submitted calls are asynchronous and run at the PCIe rate
  struct Pinned {
    uint8_t* p; uint64_t bytes;
    explicit Pinned(uint64_t b) : p(static_cast<uint8_t*>(kzg_host_alloc(b))), bytes(b) {}
    ~Pinned() { kzg_host_free(p); }
    uint8_t* data() const { return p; }
    uint8_t& operator[](uint64_t i) const { return p[i]; }
  } sc(n * 32);
  if (!sc.p) { printf("FAIL kzg_host_alloc\n"); return 1; }
  for (uint64_t i = 0; i < n; i++) scalar_u(0xC3, i, &sc[i * 32]);

  kzg_setup *s1 = nullptr, *sg = nullptr;
  OK(one, kzg_setup_generate(one, secret, 0, n, &s1));
  OK(grp, kzg_ctx_set_option(grp, "setup_layout", 1));                            // range sharding whatever n is
  OK(grp, kzg_setup_generate(grp, secret, 0, n, &sg));
  EXPECT(kzg_setup_len(sg) == n, "group setup length");
  // the shards carry the same items as the single-device setup, across every shard boundary
  {
    const uint64_t G = devs.size(), per = n / G;
    uint8_t a[96], b[96];
    for (uint64_t g = 0; g <= G; g++) {
      uint64_t first = g * per ? g * per - 1 : 0, cnt = (g == G || first + 2 > n) ? 1 : 2;
      if (first + cnt > n) first = n - cnt;
      OK(one, kzg_setup_export_g1(one, s1, first, cnt, a));
      OK(grp, kzg_setup_export_g1(grp, sg, first, cnt, b));
      EXPECT(memcmp(a, b, 48 * cnt) == 0, "sharded setup items differ from the single-device items");
    }
  }
  OK(one, kzg_setup_prepare_msm(one, s1, 0));
  OK(grp, kzg_setup_prepare_msm(grp, sg, 0));

  uint8_t c1[48], cg[48];
  OK(one, kzg_commit(one, s1, sc.data(), n, c1));
  OK(grp, kzg_commit(grp, sg, sc.data(), n, cg));
  EXPECT(memcmp(c1, cg, 48) == 0, "multi-device commitment differs from the single-device one");
  printf("commit  1 GPU %s\ncommit %2d GPU %s\n", hex(c1, 48).c_str(), (int)devs.size(), hex(cg, 48).c_str());
  if (logn == 20)   // tests/golden/bench_commitment.json (known-secret identity of the bench workload)
    EXPECT(hex(cg, 48) == "a9a29f4f21d2c6481dc286c676d58f57b1796df6e04854503c6d5348c202851fd361a5ec4624bee2e1a27092e80f7241", "golden C3 commitment");

  // shorter polynomials: one that ends inside member 0's range, one that ends inside the last range, the empty one
  for (uint64_t m : {(uint64_t)0, (uint64_t)1, n / devs.size() / 2 + 3, n - 5}) {
    OK(one, kzg_commit(one, s1, sc.data(), m, c1));
    OK(grp, kzg_commit(grp, sg, sc.data(), m, cg));
    EXPECT(memcmp(c1, cg, 48) == 0, "shorter commitment differs");
  }
  // calls in flight (tickets), 11 of them 4 deep
  {
    int32_t q[4]; int head = 0, count = 0; uint8_t out[48];
    OK(one, kzg_commit(one, s1, sc.data(), n, c1));
    auto t0 = std::chrono::steady_clock::now();
    const int calls = 11;
    for (int k = 0; k < calls; k++) {
      if (count == 4) { OK(grp, kzg_msm_g1_wait(grp, q[head], out)); EXPECT(memcmp(out, c1, 48) == 0, "pipelined result"); head = (head + 1) % 4; count--; }
      OK(grp, kzg_msm_g1_submit(grp, sg, sc.data(), n, &q[(head + count) % 4])); count++;
    }
    while (count) { OK(grp, kzg_msm_g1_wait(grp, q[head], out)); EXPECT(memcmp(out, c1, 48) == 0, "pipelined result"); head = (head + 1) % 4; count--; }
    double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    printf("%d commits of 2^%d from page-locked host memory, 4 in flight: %.2f ms each (%.1f M points/s)\n", calls, logn, ms / calls, n / (ms / calls) / 1e3);
  }
  // open: y, proof and the wrong-result error
  {
    uint8_t z[32], y1[32], yg[32], p1[48], pg[48];
    scalar_u(0xC2F, 0, z);
    OK(one, kzg_open(one, s1, sc.data(), n, z, nullptr, y1, p1));
    OK(grp, kzg_open(grp, sg, sc.data(), n, z, nullptr, yg, pg));
    EXPECT(memcmp(y1, yg, 32) == 0 && memcmp(p1, pg, 48) == 0, "open differs");
    uint8_t bad[32]; memcpy(bad, y1, 32); bad[0] ^= 1;
    EXPECT(kzg_open(grp, sg, sc.data(), n, z, bad, yg, pg) == KZG_E_REMAINDER_NONZERO, "wrong evaluation accepted");
    EXPECT(std::string(kzg_last_error(grp)).find("constant terms do not add up") != std::string::npos, "error message");
    EXPECT(kzg_commit(grp, sg, sc.data(), 0, cg) == KZG_OK && cg[0] == 0xc0, "zero polynomial -> infinity");
  }
  kzg_setup_destroy(sg); kzg_setup_destroy(s1);

  // replicated setup: a blob batch split by row (BASELINE config C4 shape, small)
  {
    const uint64_t bn = 256, blobs = 37;
    OK(grp, kzg_ctx_set_option(grp, "setup_layout", 0));
    kzg_setup *b1 = nullptr, *bg = nullptr;
    OK(one, kzg_setup_generate(one, secret, 0, bn, &b1));
    OK(grp, kzg_setup_generate(grp, secret, 0, bn, &bg));
    OK(one, kzg_setup_prepare_fixed(one, b1, bn, 8));
    OK(grp, kzg_setup_prepare_fixed(grp, bg, bn, 8));
    std::vector<uint8_t> co(blobs * bn * 32), zs(blobs * 32);
    for (uint64_t i = 0; i < blobs * bn; i++) scalar_u(0xC4, i, &co[i * 32]);
    for (uint64_t i = 0; i < blobs; i++) scalar_u(0xC4F, i, &zs[i * 32]);
    std::vector<uint8_t> cm1(blobs * 48), y1(blobs * 32), pr1(blobs * 48), cmg(blobs * 48), yg(blobs * 32), prg(blobs * 48);
    OK(one, kzg_commit_open_batch(one, b1, co.data(), blobs, bn, zs.data(), cm1.data(), y1.data(), pr1.data()));
    OK(grp, kzg_commit_open_batch(grp, bg, co.data(), blobs, bn, zs.data(), cmg.data(), yg.data(), prg.data()));
    EXPECT(cm1 == cmg && y1 == yg && pr1 == prg, "row-sharded blob batch differs");
    // a single commitment on the replicated setup runs on member 0
    OK(one, kzg_commit(one, b1, co.data(), bn, c1));
    OK(grp, kzg_commit(grp, bg, co.data(), bn, cg));
    EXPECT(memcmp(c1, cg, 48) == 0 && memcmp(cg, cmg.data(), 48) == 0, "commit on a replicated setup");
    kzg_setup_destroy(bg); kzg_setup_destroy(b1);
  }
  kzg_ctx_destroy(grp); kzg_ctx_destroy(one);
  printf("multi-device C ABI test %s\n", fails ? "FAILED" : "ok");
  return fails;
}
