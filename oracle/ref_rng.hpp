// ORACLE — TEST INFRASTRUCTURE ONLY (see ref_math.hpp header).
//
// The reference draws every random number from an entropy-seeded generator
// (`SmallRng::from_rng(thread_rng())`, implementations/src/utility/mod.rs:41-44), so its streams are
// unpinned by construction and only the *distributions* are part of its behaviour. Oracle and device
// share ONE counter-based generator instead, so that both render the same sample set:
//
//   Philox4x32-10 (Salmon et al., SC'11; Random123), key = (seed_lo, seed_hi),
//   counter = (pixel index, absolute sample index, (depth << 8) | purpose, block)
//   uniform f32 = (u32 >> 8) * 2^-24   in [0,1)   (what rand 0.8's Standard f32 does)
//
// Within one (pixel, sample, depth, purpose) stream draws are consumed sequentially: draw k is word k%4
// of block k/4.
#pragma once
#include <cstdint>

namespace ref {

struct Philox {
  static inline void mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    hi = (uint32_t)(p >> 32);
    lo = (uint32_t)p;
  }
  static inline void block(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
    uint32_t k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
      uint32_t hi0, lo0, hi1, lo1;
      mulhilo(0xD2511F53u, c0, hi0, lo0);
      mulhilo(0xCD9E8D57u, c2, hi1, lo1);
      uint32_t n0 = hi1 ^ c1 ^ k0;
      uint32_t n1 = lo1;
      uint32_t n2 = hi0 ^ c3 ^ k1;
      uint32_t n3 = lo0;
      c0 = n0; c1 = n1; c2 = n2; c3 = n3;
      k0 += 0x9E3779B9u;
      k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
  }
};

enum Purpose : uint32_t { RNG_JITTER = 0, RNG_NEE = 1, RNG_SCATTER = 2, RNG_RR = 3, RNG_TEST = 15 };

static inline float u32_to_unit(uint32_t u) { return (float)(u >> 8) * (1.0f / 16777216.0f); }

struct RngCursor {
  uint32_t key[2] = {0, 0};
  uint32_t pixel = 0, sample = 0, stream = 0;
  uint32_t drawn = 0;
  uint32_t cache[4];
  uint32_t cached_block = 0xFFFFFFFFu;
  void seed(uint64_t s) { key[0] = (uint32_t)s; key[1] = (uint32_t)(s >> 32); cached_block = 0xFFFFFFFFu; }
  void path(uint32_t pixel_, uint32_t sample_) { pixel = pixel_; sample = sample_; cached_block = 0xFFFFFFFFu; }
  void select(uint32_t depth, uint32_t purpose) {
    stream = (depth << 8) | purpose;
    drawn = 0;
    cached_block = 0xFFFFFFFFu;
  }
  uint32_t next_u32() {
    uint32_t blk = drawn >> 2;
    if (blk != cached_block) {
      uint32_t ctr[4] = {pixel, sample, stream, blk};
      Philox::block(ctr, key, cache);
      cached_block = blk;
    }
    return cache[(drawn++) & 3u];
  }
  float next_float01() { return u32_to_unit(next_u32()); }
  // gen_range(-1.0..1.0) — uniform in [-1,1)
  float next_pm1() { return 2.0f * next_float01() - 1.0f; }
  // gen_range(0..n) / gen_range(0..=n-1): uniform integer (multiply-shift)
  uint32_t next_below(uint32_t n) { return (uint32_t)(((uint64_t)next_u32() * (uint64_t)n) >> 32); }
};

// One cursor per worker thread; the integrator selects (depth, purpose) before every reference call site
// that would have created a fresh SmallRng.
extern thread_local RngCursor g_rng;

}  // namespace ref
