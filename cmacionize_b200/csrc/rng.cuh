/*
 * rng.cuh — counter-based random numbers for photon packets (Philox4x32-10,
 * Salmon et al. 2011).  Replaces the reference's per-thread ranlxd2 streams
 * (/root/reference/src/RandomGenerator.hpp:39-272): the reference's own results
 * already depend on the thread count and job schedule, so only statistical
 * parity is defined for self-generated packets (SURVEY.md §7 hard part 4).
 *
 * Stream layout: key = (seed, iteration); counter = (packet_id lo, hi, block, 0).
 * Every packet owns an independent stream that does not depend on which GPU or
 * thread processes it, so an N-GPU run draws exactly the same packets as a
 * 1-GPU run.  One Philox block yields two 53-bit uniforms in (0,1).
 */
#pragma once
#include "cmib_common.cuh"

namespace cmib {

struct PacketRng {
  uint32_t k0, k1;    /* key */
  uint32_t c0, c1;    /* packet id */
  uint32_t block;     /* block counter */
  uint32_t have;      /* 1 when `spare` holds an unused uniform */
  double spare;
};

CMIB_HD void philox_round(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, uint32_t k0,
                          uint32_t k1) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
  const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
  const uint32_t n1 = (uint32_t)p1;
  const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
  const uint32_t n3 = (uint32_t)p0;
  c0 = n0; c1 = n1; c2 = n2; c3 = n3;
}

CMIB_HD void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    philox_round(c[0], c[1], c[2], c[3], k0, k1);
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}

CMIB_HD void rng_init(PacketRng &r, uint64_t seed, uint32_t iteration, uint64_t packet_id) {
  r.k0 = (uint32_t)seed ^ (uint32_t)(seed >> 32) * 0x85EBCA6Bu;
  r.k1 = iteration * 0x9E3779B1u + (uint32_t)(seed >> 32);
  r.c0 = (uint32_t)packet_id;
  r.c1 = (uint32_t)(packet_id >> 32);
  r.block = 0;
  r.have = 0;
  r.spare = 0.;
}

CMIB_HD double u53(uint32_t hi, uint32_t lo) {
  /* 53 random bits, centred: (k + 0.5) * 2^-53 lies strictly inside (0,1) */
  const uint64_t k = (((uint64_t)hi << 32) | lo) >> 11;
  return ((double)k + 0.5) * (1.0 / 9007199254740992.0);
}

CMIB_HD double rng_uniform(PacketRng &r) {
  if (r.have) {
    r.have = 0;
    return r.spare;
  }
  uint32_t c[4] = {r.c0, r.c1, r.block, 0u};
  philox4x32_10(c, r.k0, r.k1);
  ++r.block;
  r.spare = u53(c[2], c[3]);
  r.have = 1;
  return u53(c[0], c[1]);
}

} // namespace cmib
