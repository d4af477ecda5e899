// Host build of the engine's recoder (recode.cuh is host + device: the same SignedDigits the kernels use): prints, for every scalar on stdin (hex,
// little-endian value < 2^256) and every window size on the command line, the folded scalar and its signed digits
// as msm.cu k_recode, batch.cu, batch_tree.cu digits_of, setup.cu and setup_g2.cu compute them:
//   <c> <flip> <W> d_0 d_1 ... d_{W-1} <final carry>          (d_j signed, sign already combined with the fold)
// tests/test_recode_cpu.py checks sum d_j 2^(c j) = k (mod r), |d_j| <= 2^(c-1) and final carry = 0 against Python bigints.
// Build: g++ -std=c++17 -O1 -I kzg_poly_commit_exploration_b200/csrc -x c++ tests/cpp/test_recode.cpp -o tests/cpp/test_recode.bin
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <iostream>
#include "recode.cuh"
using namespace kzg;
int main(int argc, char** argv) {
  std::string line;
  while (std::getline(std::cin, line)) {
    if (line.empty()) continue;
    Fr k0; memset(&k0, 0, sizeof k0);
    std::string h = line;                                   // hex, most significant digit first
    while (h.size() < 64) h = "0" + h;
    for (int i = 0; i < 8; i++) k0.v[i] = (uint32_t)strtoul(h.substr(64 - 8 * (i + 1), 8).c_str(), nullptr, 16);
    fr_canon_reduce(k0);
    for (int a = 1; a < argc; a++) {
      const int c = atoi(argv[a]), W = windows_for(c);
      SignedDigits sd; sd.init(k0, c);
      printf("%d %d %d", c, sd.flip ? 1 : 0, W);
      for (int j = 0; j < W; j++) {
        bool neg;
        const uint32_t d = sd.next(j, &neg);
        printf(" %lld", neg ? -(long long)d : (long long)d);
      }
      printf(" %u\n", sd.carry);
    }
  }
  return 0;
}
