// Signed-digit window recoding of a scalar, shared by every kernel that walks a table (host + device so that the CPU
// test-suite runs the very same code: tests/cpp/test_recode.cpp).
//
// A canonical scalar k < r is first FOLDED to min(k, r - k) (fr_fold_half, field.cuh; [k]P = -[r-k]P flips the sign of
// every digit), which leaves k < 2^254.  Then digit_j = ((k >> c j) & (2^c - 1)) + carry; a digit above 2^(c-1) becomes
// digit - 2^c with a carry into the next window.  With W = ceil(255 / c) windows the top window holds at
// most 254 - c (W - 1) <= c - 1 bits, so digit + carry <= 2^(c-1) there and no carry leaves it: sum_j d_j 2^(c j) = +-k.
#pragma once
#include "field.cuh"

namespace kzg {

// windows needed for a folded scalar (c = 16 -> 16, c = 15 -> 17, c = 17 -> 15, c = 13 -> 20)
KZG_HD int windows_for(int c) { return (255 + c - 1) / c; }

// bits [bit, bit + c) of a 256-bit integer in 32-bit limbs (zero beyond bit 255), c <= 24
KZG_HD uint32_t window_bits(const uint32_t* k, int bit, int c) {
  const int w = bit >> 5, sh = bit & 31;
  if (w >= 8) return 0;
  uint64_t lo = k[w];
  if (w + 1 < 8) lo |= (uint64_t)k[w + 1] << 32;
  return (uint32_t)(lo >> sh) & ((1u << c) - 1u);
}

struct SignedDigits {
  Fr k;              // the folded scalar
  bool flip;         // k was replaced by r - k: every digit changes sign
  uint32_t carry;
  int c;
  // `scalar` canonical (< r); fold = false keeps the scalar as it is (callers whose window count is ceil(256 / c) anyway)
  KZG_HD void init(const Fr& scalar, int window_bits_, bool fold = true) {
    k = scalar; c = window_bits_; carry = 0;
    flip = fold ? fr_fold_half(k) : false;
  }
  // digit of window j (call with j = 0, 1, 2, ... in order): magnitude in [0, 2^(c-1)], *neg = its sign (fold included)
  KZG_HD uint32_t next(int j, bool* neg) {
    uint32_t d = window_bits(k.v, j * c, c) + carry;
    const bool negd = d > (1u << (c - 1));
    if (negd) { d = (1u << c) - d; carry = 1; } else carry = 0;
    *neg = negd != flip;
    return d;
  }
};

}  // namespace kzg
