// dart:math `Random(seed)` as the Dart VM implements it, for the one place on the GPU path that needs the reference's own
// generator rather than a keyed stream: BestCandidateSampler seeds a fresh RNG per table tile with xTile + (yTile << 8)
// (lib/samplers/best_candidate_sampler.dart:44-47,91-94; lib/core/rng.dart:27-43) and draws the tile's three sample shifts.
// The algorithm lives in the Dart SDK (runtime/lib/math_patch.dart), not in the reference tree, and the SDK version is not
// pinned by the reference (no `environment:` in pubspec.yaml): a 64-bit multiply-with-carry state seeded through a 64-bit mix.
#pragma once
#include <cstdint>

namespace drt {

class DartRandom {
 public:
  explicit DartRandom(int64_t seed) {
    uint64_t n = (uint64_t)seed;
    n = (~n) + (n << 21);
    n ^= n >> 24;
    n *= 265;
    n ^= n >> 14;
    n *= 21;
    n ^= n >> 28;
    n += n << 31;
    if (n == 0) n = 0x5A17;
    lo_ = (uint32_t)n;
    hi_ = (uint32_t)(n >> 32);
    for (int i = 0; i < 4; ++i) step();
  }
  // nextDouble(): 26 + 27 random bits over 2^53
  double nextDouble() {
    const double a = (double)bits(26), b = (double)bits(27);
    return (a * 134217728.0 + b) / 9007199254740992.0;
  }

 private:
  void step() {
    const uint64_t s = 0xffffda61ull * lo_ + hi_;
    lo_ = (uint32_t)s;
    hi_ = (uint32_t)(s >> 32);
  }
  uint32_t bits(int k) {  // nextInt(1 << k): a power of two takes the low bits of the next state
    step();
    return lo_ & ((1u << k) - 1u);
  }
  uint32_t lo_, hi_;
};

}  // namespace drt
