// Host build of the engine's recoding helpers (field.cuh is host + device): prints, for every scalar on stdin (hex,
// little-endian value < 2^256) and every window size on the command line, the folded scalar and its signed digits
// exactly as the CUDA recoders (msm.cu k_recode, batch_tree.cu digits_of, setup_g2.cu g2_digits) compute them:
//   <c> <flip> <W> d_0 d_1 ... d_{W-1} <final carry>          (d_j signed, sign already combined with the fold)
// tests/test_recode_cpu.py checks sum d_j 2^(c j) = k (mod r), |d_j| <= 2^(c-1) and final carry = 0 against Python bigints.
// Build: g++ -std=c++17 -O1 -I kzg_poly_commit_exploration_b200/csrc -x c++ tests/cpp/test_recode.cpp -o tests/cpp/test_recode.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <iostream>
#include "field.cuh"
using namespace kzg;
static inline int windows_for(int c) { return (255 + c - 1) / c; }          // engine.cuh
int main(int argc, char** argv) {
  std::string line;
  while (std::getline(std::cin, line)) {
    if (line.empty()) continue;
    Fr k0; memset(&k0, 0, sizeof k0);
    // parse hex big-endian string into 8 little-endian limbs
    std::string h = line;
    while (h.size() < 64) h = "0" + h;
    for (int i = 0; i < 8; i++) k0.v[i] = (uint32_t)strtoul(h.substr(64 - 8 * (i + 1), 8).c_str(), nullptr, 16);
    fr_canon_reduce(k0);
    for (int a = 1; a < argc; a++) {
      const int c = atoi(argv[a]), W = windows_for(c);
      Fr k = k0;
      const bool flip = fr_fold_half(k);
      const uint32_t half = 1u << (c - 1);
      uint32_t carry = 0;
      printf("%d %d %d", c, flip ? 1 : 0, W);
      for (int j = 0; j < W; j++) {
        const int bit = j * c, w = bit >> 5, sh = bit & 31;
        uint32_t d = 0;
        if (w < 8) {
          uint64_t lo = k.v[w];
          if (w + 1 < 8) lo |= (uint64_t)k.v[w + 1] << 32;
          d = (uint32_t)(lo >> sh) & ((1u << c) - 1u);
        }
        d += carry;
        const bool negd = d > half;
        if (negd) { d = (1u << c) - d; carry = 1; } else carry = 0;
        long long sd = (negd != flip) ? -(long long)d : (long long)d;
        printf(" %lld", sd);
      }
      printf(" %u\n", carry);
    }
  }
  return 0;
}
