/*
 * linecooling.cuh — collisionally excited line cooling by 10 five-level and 3
 * two-level metal ions.
 *
 * Behavioural contract:
 *   LineCoolingData::get_cooling               /root/reference/src/LineCoolingData.cpp:1767-1848
 *   LineCoolingData::compute_level_populations ...:1569-1700 (5x5 rate matrix)
 *   LineCoolingData::solve_system_of_linear_equations ...:1492-1555
 *   LineCoolingData::compute_level_population  ...:1714-1736 (two-level closed form)
 *   element order (NI NII OI OII OIII NeIII SII SIII CII CIII | NIII NeII SIV)
 *                                              LineCoolingData.hpp:38-75
 *
 * Design for the GPU: the atomic data live in one 7.9 KB __constant__ table
 * (all lanes walk elements/transitions in lock step, so every read is a
 * constant-cache broadcast); the 5x5 system lives entirely in registers with
 * fully unrolled loops.  The reference always finds the solution by partial
 * pivoting; pivot selection is data dependent, so the elimination is written
 * with predicated row swaps (no dynamic register indexing).
 */
#pragma once
#include "cmib_common.cuh"
#include "tables.cuh"

namespace cmib {

enum LineCoolElement : int {
  LC_NI = 0, LC_NII, LC_OI, LC_OII, LC_OIII, LC_NeIII, LC_SII, LC_SIII, LC_CII, LC_CIII,
  LC_NUM5,
  LC_NIII = LC_NUM5, LC_NeII, LC_SIV, LC_NUM
};

/* offsets into the flat LINECOOLING table (tools/gen_linecooling_data.py) */
constexpr int LC_OFF_CS5 = 0;
constexpr int LC_OFF_A5 = LC_OFF_CS5 + 10 * 10 * 7;
constexpr int LC_OFF_E5 = LC_OFF_A5 + 100;
constexpr int LC_OFF_W5 = LC_OFF_E5 + 100;
constexpr int LC_OFF_CS2 = LC_OFF_W5 + 50;
constexpr int LC_OFF_A2 = LC_OFF_CS2 + 21;
constexpr int LC_OFF_E2 = LC_OFF_A2 + 3;
constexpr int LC_OFF_W2 = LC_OFF_E2 + 3;
constexpr int LC_OFF_PREFACTOR = LC_OFF_W2 + 6;

/* transitions: 0:0-1 1:0-2 2:0-3 3:0-4 4:1-2 5:1-3 6:1-4 7:2-3 8:2-4 9:3-4 */

/* Gaussian elimination with partial pivoting on a 5x5 system, same pivoting rule
 * and operation order as the reference; returns 1 if singular. */
CMIB_HD int solve5(double A[5][5], double B[5]) {
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    int imax = 0;
    double Amax = 0.;
#pragma unroll
    for (int i = j; i < 5; ++i) {
      if (fabs(A[i][j]) > fabs(Amax)) {
        Amax = A[i][j];
        imax = i;
      }
    }
    if (Amax == 0.) return 1;
    const double Amax_inv = 1. / Amax;
    /* predicated swap of row j with row imax (imax >= j) */
#pragma unroll
    for (int i = j + 1; i < 5; ++i) {
      if (imax == i) {
#pragma unroll
        for (int k = 0; k < 5; ++k) {
          const double save = A[j][k];
          A[j][k] = A[i][k];
          A[i][k] = save;
        }
        const double save = B[j];
        B[j] = B[i];
        B[i] = save;
      }
    }
#pragma unroll
    for (int k = 0; k < 5; ++k) A[j][k] *= Amax_inv;
    B[j] *= Amax_inv;
    if (j < 4) {
#pragma unroll
      for (int i = j + 1; i < 5; ++i) {
#pragma unroll
        for (int k = j + 1; k < 5; ++k) A[i][k] -= A[i][j] * A[j][k];
        B[i] -= A[i][j] * B[j];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < i + 1; ++j) B[3 - i] -= B[4 - j] * A[3 - i][4 - j];
  }
  return 0;
}

/* where the table is read from: __constant__ memory when all lanes of a warp evaluate the same ion at the same time (the
 * constant-cache broadcast), a plain pointer (shared memory copy) when the lanes of a warp evaluate DIFFERENT ions
 * (line_cooling_wide of kernels.cuh) */
struct LcTableConst {
  CMIB_HD double operator[](int i) const { return CMIB_TBL(LINECOOLING)[i]; }
};
struct LcTablePtr {
  const double *p;
  CMIB_HD double operator[](int i) const { return p[i]; }
};

/* c = offset of the 7 fit coefficients of the transition */
template <class Tab>
CMIB_HD double collision_strength(const Tab &tab, int c, double prefactor, double T, double Tinv, double logT) {
  return prefactor * powl(T, logT, 1. + tab[c]) *
         (tab[c + 1] + tab[c + 2] * Tinv + tab[c + 3] * logT + tab[c + 4] * T * (1. + (tab[c + 5] - 1.) * powl(T, logT, tab[c + 6])));
}

/* level populations of five-level element e; returns solver status */
template <class Tab>
CMIB_HD int five_level_populations(const Tab &tab, int e, double prefactor, double T, double Tinv, double logT,
                                   double pop[5]) {
  const int oA = LC_OFF_A5 + 10 * e, oE = LC_OFF_E5 + 10 * e, ow = LC_OFF_W5 + 5 * e;
  double A[10], w[5];
#pragma unroll
  for (int t = 0; t < 10; ++t) A[t] = tab[oA + t];
#pragma unroll
  for (int t = 0; t < 5; ++t) w[t] = tab[ow + t];
  double dn[10], up[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    const double cs = collision_strength(tab, LC_OFF_CS5 + (e * 10 + t) * 7, prefactor, T, Tinv, logT);
    dn[t] = cs;
    up[t] = cs * exp(-tab[oE + t] * Tinv);
  }
  double M[5][5];
#pragma unroll
  for (int i = 0; i < 5; ++i) {
    M[0][i] = 1.;
    pop[i] = 0.;
  }
  pop[0] = 1.;
  M[1][0] = up[0] * w[0];
  M[1][1] = -(A[0] + w[1] * (dn[0] + up[4] + up[5] + up[6]));
  M[1][2] = A[4] + w[2] * dn[4];
  M[1][3] = A[5] + w[3] * dn[5];
  M[1][4] = A[6] + w[4] * dn[6];
  M[2][0] = up[1] * w[0];
  M[2][1] = up[4] * w[1];
  M[2][2] = -(A[1] + A[4] + w[2] * (dn[1] + dn[4] + up[7] + up[8]));
  M[2][3] = A[7] + dn[7] * w[3];
  M[2][4] = A[8] + dn[8] * w[4];
  M[3][0] = up[2] * w[0];
  M[3][1] = up[5] * w[1];
  M[3][2] = up[7] * w[2];
  M[3][3] = -(A[2] + A[5] + A[7] + w[3] * (dn[2] + dn[5] + dn[7] + up[9]));
  M[3][4] = A[9] + dn[9] * w[4];
  M[4][0] = up[3] * w[0];
  M[4][1] = up[6] * w[1];
  M[4][2] = up[8] * w[2];
  M[4][3] = up[9] * w[3];
  M[4][4] = -(A[3] + A[6] + A[8] + A[9] + w[4] * (dn[3] + dn[6] + dn[8] + dn[9]));
  return solve5(M, pop);
}
CMIB_HD int five_level_populations(int e, double prefactor, double T, double Tinv, double logT, double pop[5]) {
  return five_level_populations(LcTableConst(), e, prefactor, T, Tinv, logT, pop);
}

template <class Tab>
CMIB_HD double two_level_population(const Tab &tab, int i, double prefactor, double T, double Tinv, double logT) {
  const double ksi = tab[LC_OFF_E2 + i];
  const double A = tab[LC_OFF_A2 + i];
  const double cs = collision_strength(tab, LC_OFF_CS2 + 7 * i, prefactor, T, Tinv, logT);
  const double inv_omega_1 = tab[LC_OFF_W2 + 2 * i];
  const double inv_omega_2 = tab[LC_OFF_W2 + 2 * i + 1];
  const double Texp = exp(-ksi * Tinv);
  return cs * Texp * inv_omega_1 / (A + cs * (inv_omega_2 + Texp * inv_omega_1));
}
CMIB_HD double two_level_population(int i, double prefactor, double T, double Tinv, double logT) {
  return two_level_population(LcTableConst(), i, prefactor, T, Tinv, logT);
}

/* the terms of the cooling sum (LineCoolingData::get_cooling, LineCoolingData.cpp): five-level element e, two-level
 * element i; abund = the abundance of that element's ion */
template <class Tab>
CMIB_HD double line_cooling_term5(const Tab &tab, int e, double prefactor, double T, double Tinv, double logT, double abund) {
  double pop[5];
  five_level_populations(tab, e, prefactor, T, Tinv, logT, pop);
  const int oA = LC_OFF_A5 + 10 * e, oE = LC_OFF_E5 + 10 * e;
  const double cl2 = pop[1] * tab[oA] * tab[oE];
  const double cl3 = pop[2] * (tab[oA + 1] * tab[oE + 1] + tab[oA + 4] * tab[oE + 4]);
  const double cl4 = pop[3] * (tab[oA + 2] * tab[oE + 2] + tab[oA + 5] * tab[oE + 5] + tab[oA + 7] * tab[oE + 7]);
  const double cl5 = pop[4] * (tab[oA + 3] * tab[oE + 3] + tab[oA + 6] * tab[oE + 6] + tab[oA + 8] * tab[oE + 8] + tab[oA + 9] * tab[oE + 9]);
  return abund * BOLTZMANN * (cl2 + cl3 + cl4 + cl5);
}
template <class Tab>
CMIB_HD double line_cooling_term2(const Tab &tab, int i, double prefactor, double T, double Tinv, double logT, double abund) {
  const double lp = two_level_population(tab, i, prefactor, T, Tinv, logT);
  return abund * BOLTZMANN * tab[LC_OFF_E2 + i] * tab[LC_OFF_A2 + i] * lp;
}

/* cooling rate per hydrogen atom (J s^-1); abund in LineCoolElement order */
CMIB_HD double line_cooling(double T, double ne, const double abund[LC_NUM]) {
  if (ne == 0.) return 1.e-99;
  const LcTableConst tab;
  const double prefactor = tab[LC_OFF_PREFACTOR] * ne / sqrt(T);
  const double Tinv = 1. / T;
  const double logT = log(T);
  double cooling = 0.;
  for (int e = 0; e < LC_NUM5; ++e) cooling += line_cooling_term5(tab, e, prefactor, T, Tinv, logT, abund[e]);
  for (int i = 0; i < 3; ++i) cooling += line_cooling_term2(tab, i, prefactor, T, Tinv, logT, abund[LC_NUM5 + i]);
  return cooling;
}

constexpr int LC_TABLE_SIZE = LC_OFF_PREFACTOR + 1;

} // namespace cmib
