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

CMIB_HD double collision_strength(const double *c, double prefactor, double T, double Tinv,
                                  double logT) {
  return prefactor * powl(T, logT, 1. + c[0]) *
         (c[1] + c[2] * Tinv + c[3] * logT + c[4] * T * (1. + (c[5] - 1.) * powl(T, logT, c[6])));
}

/* level populations of five-level element e; returns solver status */
CMIB_HD int five_level_populations(int e, double prefactor, double T, double Tinv, double logT,
                                   double pop[5]) {
  const double *tab = CMIB_TBL(LINECOOLING);
  const double *A = tab + LC_OFF_A5 + 10 * e;
  const double *E = tab + LC_OFF_E5 + 10 * e;
  const double *w = tab + LC_OFF_W5 + 5 * e;
  double dn[10], up[10];
#pragma unroll
  for (int t = 0; t < 10; ++t) {
    const double cs = collision_strength(tab + LC_OFF_CS5 + (e * 10 + t) * 7, prefactor, T, Tinv, logT);
    dn[t] = cs;
    up[t] = cs * exp(-E[t] * Tinv);
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

CMIB_HD double two_level_population(int i, double prefactor, double T, double Tinv, double logT) {
  const double *tab = CMIB_TBL(LINECOOLING);
  const double ksi = tab[LC_OFF_E2 + i];
  const double A = tab[LC_OFF_A2 + i];
  const double cs = collision_strength(tab + LC_OFF_CS2 + 7 * i, prefactor, T, Tinv, logT);
  const double inv_omega_1 = tab[LC_OFF_W2 + 2 * i];
  const double inv_omega_2 = tab[LC_OFF_W2 + 2 * i + 1];
  const double Texp = exp(-ksi * Tinv);
  return cs * Texp * inv_omega_1 / (A + cs * (inv_omega_2 + Texp * inv_omega_1));
}

/* cooling rate per hydrogen atom (J s^-1); abund in LineCoolElement order */
CMIB_HD double line_cooling(double T, double ne, const double abund[LC_NUM]) {
  if (ne == 0.) return 1.e-99;
  const double *tab = CMIB_TBL(LINECOOLING);
  const double prefactor = tab[LC_OFF_PREFACTOR] * ne / sqrt(T);
  const double Tinv = 1. / T;
  const double logT = log(T);
  double cooling = 0.;
  for (int e = 0; e < LC_NUM5; ++e) {
    double pop[5];
    five_level_populations(e, prefactor, T, Tinv, logT, pop);
    const double *A = tab + LC_OFF_A5 + 10 * e;
    const double *E = tab + LC_OFF_E5 + 10 * e;
    const double cl2 = pop[1] * A[0] * E[0];
    const double cl3 = pop[2] * (A[1] * E[1] + A[4] * E[4]);
    const double cl4 = pop[3] * (A[2] * E[2] + A[5] * E[5] + A[7] * E[7]);
    const double cl5 = pop[4] * (A[3] * E[3] + A[6] * E[6] + A[8] * E[8] + A[9] * E[9]);
    cooling += abund[e] * BOLTZMANN * (cl2 + cl3 + cl4 + cl5);
  }
  for (int i = 0; i < 3; ++i) {
    const double lp = two_level_population(i, prefactor, T, Tinv, logT);
    cooling += abund[LC_NUM5 + i] * BOLTZMANN * tab[LC_OFF_E2 + i] * tab[LC_OFF_A2 + i] * lp;
  }
  return cooling;
}

} // namespace cmib
