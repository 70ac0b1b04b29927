/*
 * RandomGenerator.hpp — the reference's host-side random stream (RANLUX, double precision,
 * luxury level 2: Luescher 1994), for the plugins whose OUTPUT depends on it: the positions a
 * UniformRandomPhotonSourceDistribution draws must be the ones the reference draws for the
 * same seed, or the two codes simulate different problems.
 *
 * Behavioural contract = RandomGenerator (/root/reference/src/RandomGenerator.hpp:83-225):
 * same seeding (a 31-bit shift register fills twelve 48-bit fractions), same
 * subtract-with-borrow recurrence x[i] <- x[i+7] - x[i] - carry (mod 1) on a ring of 12,
 * 397 updates per 12 numbers delivered.  All values are multiples of 2^-48, so the arithmetic
 * is exact and the stream is bit-identical (tests/test_host_layer.py checks 10^5 deviates
 * for several seeds against the compiled reference).
 *
 * The photon packets do NOT use this generator: they draw from the counter-based Philox
 * streams of csrc/rng.cuh (one independent stream per packet id, which is what lets the shoot
 * shard over GPUs); only statistical parity is claimed there.
 */
#pragma once
#include <cstdint>

namespace cmi {

class RandomGenerator {
public:
  explicit RandomGenerator(int32_t seed = 42) { set_seed(seed); }

  void set_seed(int32_t seed) {
    if (seed == 0) seed = 1;
    /* 31 seed bits drive the linear feedback shift register b[n] = b[n-31] ^ b[n-13] */
    int bit[31];
    uint32_t v = (uint32_t)seed & 0x7fffffffu;
    for (int k = 0; k < 31; ++k, v >>= 1) bit[k] = (int)(v & 1u);
    int head = 0, tap = 18;
    for (int k = 0; k < RING; ++k) {
      double x = 0.;
      for (int m = 0; m < 48; ++m) {
        x = 2. * x + (double)(1 - bit[head]); /* complemented register output, MSB first */
        bit[head] ^= bit[tap];
        head = (head + 1) % 31;
        tap = (tap + 1) % 31;
      }
      x_[k] = ULP48 * x;
    }
    carry_ = 0.;
    next_ = RING - 1;
    partner_ = 7;
    refill_at_ = 0;
  }

  /* uniform deviate in [0, 1) */
  double get_uniform_random_double() {
    next_ = (next_ + 1) % RING;
    if (next_ == refill_at_) advance();
    return x_[next_];
  }

  int32_t get_random_integer() { return (int32_t)(get_uniform_random_double() * 2147483648.0); }

private:
  static constexpr int RING = 12;
  static constexpr int UPDATES = 397; /* luxury level 2 */
  static constexpr double ULP48 = 1.0 / 281474976710656.0;

  /* UPDATES steps of the subtract-with-borrow recurrence (the reference unrolls the middle of
   * this loop by 12 with the borrow folded into the next difference: same operations) */
  void advance() {
    int i = next_, j = partner_;
    for (int k = 0; k < UPDATES; ++k) {
      double y = (x_[j] - x_[i]) - carry_;
      if (y < 0.) {
        carry_ = ULP48;
        y += 1.;
      } else {
        carry_ = 0.;
      }
      x_[i] = y;
      i = (i + 1) % RING;
      j = (j + 1) % RING;
    }
    next_ = i;
    refill_at_ = i;
    partner_ = j;
  }

  double x_[RING];
  double carry_;
  int next_, partner_, refill_at_;
};

} // namespace cmi
