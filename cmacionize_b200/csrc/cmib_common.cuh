/*
 * cmib_common.cuh — shared types, enums and exact-arithmetic helpers for the
 * B200 photoionization hot path.
 *
 * Names follow the reference's domain vocabulary (cells, packets, ions):
 *   IonName order            /root/reference/src/ElementNames.hpp:107-160
 *   HeatingTermName          /root/reference/src/IonizationVariables.hpp:66-76
 *   ReemissionProbabilityName /root/reference/src/IonizationVariables.hpp:46-61
 *   PhotonType               /root/reference/src/PhotonType.hpp:41-56
 */
#pragma once
#include <cfloat>
#include <cmath>
#include <cstdint>

#if defined(__CUDACC__)
#define CMIB_HD __host__ __device__ __forceinline__
#define CMIB_D __device__ __forceinline__
#else
#define CMIB_HD inline
#define CMIB_D inline
#endif

#if !defined(__CUDACC__) && !defined(__VECTOR_TYPES_H__)
/* host build of the shared headers without the CUDA headers (tests/hostcheck): CUDA's vector type */
struct double2 { double x, y; };
#endif

namespace cmib {

enum Ion : int {
  ION_H_n = 0, ION_He_n, ION_C_p1, ION_C_p2, ION_N_n, ION_N_p1, ION_N_p2, ION_O_n,
  ION_O_p1, ION_Ne_n, ION_Ne_p1, ION_S_p1, ION_S_p2, ION_S_p3, NUM_IONS
};
enum Element : int { EL_He = 0, EL_C, EL_N, EL_O, EL_Ne, EL_S, NUM_ELEMENTS };
enum HeatTerm : int { HEAT_H = 0, HEAT_He, NUM_HEAT };
enum ReemitProb : int {
  REEMIT_H = 0, REEMIT_HE_LYC, REEMIT_HE_NPEEV, REEMIT_HE_TPC, REEMIT_HE_LYA, NUM_REEMIT
};
enum PacketType : int {
  PACKET_PRIMARY = 0, PACKET_DIFFUSE_HI, PACKET_DIFFUSE_HeI, PACKET_ABSORBED, NUM_PACKET_TYPES
};

/* number of accumulators per cell in the full layout: J[14] then heat[2] */
constexpr int NUM_ACC = NUM_IONS + NUM_HEAT;

/* CODATA 2014 values used by the reference (PhysicalConstants.hpp:73-110) */
constexpr double PLANCK = 6.626070040e-34;
constexpr double BOLTZMANN = 1.38064852e-23;
constexpr double ELECTRONVOLT = 1.6021766208e-19;

/*
 * Exact (non-contracted) FP64 arithmetic.  The reference is built without FMA
 * (-std=c++11 -O3, no -march), nvcc contracts a*b+c by
 * default.  Everything that decides WHICH cell a packet visits goes through
 * these so that the traversal is bit-identical to CartesianDensityGrid::interact.
 */
#if defined(__CUDA_ARCH__)
CMIB_HD double xmul(double a, double b) { return __dmul_rn(a, b); }
CMIB_HD double xadd(double a, double b) { return __dadd_rn(a, b); }
CMIB_HD double xsub(double a, double b) { return __dsub_rn(a, b); }
CMIB_HD double xdiv(double a, double b) { return __ddiv_rn(a, b); }
#else
/* host build (tests/hostcheck only): compiled with -ffp-contract=off */
CMIB_HD double xmul(double a, double b) { return a * b; }
CMIB_HD double xadd(double a, double b) { return a + b; }
CMIB_HD double xsub(double a, double b) { return a - b; }
CMIB_HD double xdiv(double a, double b) { return a / b; }
#endif

/*
 * x^a for the smooth temperature / frequency fits (recombination, charge transfer, collision
 * strengths, re-emission probabilities).  Host build: pow(), the reference's call, so the CPU tier
 * pins every fit bit for bit.  Device: exp(a ln x) — FP64 pow costs about 4 exp on sm_100a and a
 * temperature solve evaluates ~270 of these per heating/cooling balance.  |a ln x| < ~30 for every
 * fit on the path, so the value moves by a few 1e-15 relative; the GPU-tier tests state and measure
 * the resulting bound for every function.  powl(lnx, a) is the same with the logarithm supplied.
 */
CMIB_HD double fpow(double x, double a) {
#if defined(__CUDA_ARCH__)
  return exp(a * log(x));
#else
  return pow(x, a);
#endif
}
CMIB_HD double powl(double x, double lnx, double a) {
#if defined(__CUDA_ARCH__)
  (void)x;
  return exp(a * lnx);
#else
  (void)lnx;
  return pow(x, a);
#endif
}

/* geometry of the Cartesian grid; constants derived exactly as in
 * CartesianDensityGrid.cpp:80-86 (cellside = sides / ncell; inv = 1. / cellside) */
struct GridGeom {
  double anchor[3];
  double sides[3];
  double cellside[3];
  double inv_cellside[3];
  int32_t ncell[3];
  int32_t periodic[3];
  double cell_volume; /* cellside.x * cellside.y * cellside.z (CartesianDensityGrid.hpp:98-100) */
  int64_t ncells;
};

/* read-mostly per-cell record the ray march gathers: one 32-byte sector */
struct __attribute__((aligned(32))) CellOpacity {
  double n;   /* number density (m^-3) */
  double xH;  /* neutral fraction of H */
  double xHe; /* neutral fraction of He */
  double T;   /* temperature (K); used by the re-emission draw only */
};

} // namespace cmib
