// pred_i32.h -- the arithmetic of the vectorised NEXMark predicate `CAST(i32col AS Int64) [% m] CMP rhs`, shared by the
// device functor (filter_project.cu: PredI32) and the CPU-only self test (selftest.cc, tests/test_host.py), so that the
// constants and the per-row test the GPU runs are checked against Rust's `%` semantics without a GPU.
#pragma once

#include <algorithm>
#include <cstdint>

#include "expr_program.h"

namespace fg {

// Three arithmetic shapes:
//   MODE 0  affine range test  keep = (uint32(x) * mul + add <= span) != neg      -- one IMAD + one compare per row.
//           Covers every plain comparison (mul = 1, add = -lo, span = hi - lo for the interval [lo, hi] of
//           accepted values, `neg` for the complement) AND `% m (=|!=) 0` for odd m: with inv = m^-1 mod 2^32,
//           L = (2^31 - 1) / m, L2 = 2^31 / m, x is a multiple of m iff x * inv mod 2^32 lies in [0, L] (x >= 0)
//           or in [2^32 - L2, 2^32) (x < 0), i.e. iff x * inv + L2 <= L + L2 (Granlund-Montgomery exact division
//           test, Hacker's Delight 10-17, extended to signed dividends).
//   MODE 1  general `% m CMP c` via Lemire's fastmod (M = 2^64 / m + 1: two multiplies instead of a division).
//   MODE 2  `% m (=|!=) 0` for even m = 2^k q: rotr(|x| * q^-1, k) <= (2^32 - 1) / m.
struct PredI32Consts {
  // MODE 0 / 2
  uint32_t mul, add, span, rot;
  int32_t neg;
  // MODE 1
  int32_t cmp;
  uint32_t d;
  uint64_t M;
  int64_t rhs;
};

FG_HD bool pred_i32_cmp_i64(int cmp, int64_t a, int64_t b) {
  switch (cmp) {
    case FLOCKGPU_OP_EQ: return a == b;
    case FLOCKGPU_OP_NE: return a != b;
    case FLOCKGPU_OP_LT: return a < b;
    case FLOCKGPU_OP_LE: return a <= b;
    case FLOCKGPU_OP_GT: return a > b;
    default: return a >= b;
  }
}

// The per-row test of mode MODE (0 / 1 / 2, see above).
template <int MODE>
FG_HD bool pred_i32_test(const PredI32Consts& k, int32_t x) {
  if (MODE == 0) return (uint32_t(x) * k.mul + k.add <= k.span) != bool(k.neg);
  const uint32_t ax = x < 0 ? 0u - uint32_t(x) : uint32_t(x);
  if (MODE == 2) {
    const uint32_t m = ax * k.mul;
#ifdef __CUDA_ARCH__
    const uint32_t r = __funnelshift_r(m, m, k.rot);
#else
    const uint32_t r = k.rot ? ((m >> k.rot) | (m << (32 - k.rot))) : m;
#endif
    return (r <= k.span) != bool(k.neg);
  }
  const uint64_t low = k.M * uint64_t(ax);
#ifdef __CUDA_ARCH__
  int64_t r = int64_t(__umul64hi(low, uint64_t(k.d)));
#else
  int64_t r = int64_t(uint64_t((static_cast<unsigned __int128>(low) * k.d) >> 64));
#endif
  if (x < 0) r = -r;
  return pred_i32_cmp_i64(k.cmp, r, k.rhs);
}

// Host side: the constants of PredI32 for `CAST(col AS Int64) [% modulus] cmp rhs` (modulus = 0: no `%`).
// Returns the MODE to launch.
inline int pred_i32_consts(int64_t modulus, int cmp, int64_t rhs, PredI32Consts* out) {
  PredI32Consts k{};
  k.cmp = cmp;
  k.rhs = rhs;
  if (modulus == 0) {
    // accepted interval [lo, hi] of the positive form; NE is the complement of EQ
    const int64_t MIN = INT32_MIN, MAX = INT32_MAX;
    int64_t lo = MIN, hi = MAX;
    bool neg = false;
    switch (cmp) {
      case FLOCKGPU_OP_EQ: lo = hi = rhs; break;
      case FLOCKGPU_OP_NE: lo = hi = rhs; neg = true; break;
      case FLOCKGPU_OP_LT: hi = rhs > MIN ? rhs - 1 : MIN - 1; break;
      case FLOCKGPU_OP_LE: hi = rhs; break;
      case FLOCKGPU_OP_GT: lo = rhs < MAX ? rhs + 1 : MAX + 1; break;
      default: lo = rhs; break;  // GE
    }
    lo = std::max(lo, MIN);
    hi = std::min(hi, MAX);
    if (lo > hi) {  // nothing in the int32 domain: the complement of everything
      lo = MIN;
      hi = MAX;
      neg = !neg;
    }
    k.mul = 1u;
    k.add = 0u - uint32_t(int32_t(lo));
    k.span = uint32_t(hi - lo);
    k.neg = neg;
    *out = k;
    return 0;
  }
  const uint32_t d = uint32_t(modulus);
  k.d = d;
  k.M = ~uint64_t(0) / d + 1;
  const bool divisibility = rhs == 0 && (cmp == FLOCKGPU_OP_EQ || cmp == FLOCKGPU_OP_NE);
  if (!divisibility) {
    *out = k;
    return 1;
  }
  uint32_t rot = 0, q = d;
  while (!(q & 1u)) {
    q >>= 1;
    ++rot;
  }
  uint32_t inv = q;  // Newton: doubles the number of correct low bits per step (3 -> 96)
  for (int it = 0; it < 5; ++it) inv *= 2u - q * inv;
  k.mul = inv;
  k.neg = cmp == FLOCKGPU_OP_NE;
  if (rot == 0) {
    const uint32_t L = 0x7fffffffu / d, L2 = 0x80000000u / d;
    k.add = L2;
    k.span = L + L2;
    *out = k;
    return 0;
  }
  k.rot = rot;
  k.span = 0xffffffffu / d;
  *out = k;
  return 2;
}

}  // namespace fg
