/*
 * rates.cuh — recombination and charge-transfer rate coefficients.
 *
 * Behavioural contract:
 *   RecombinationRates::get_recombination_rate(ion, T)   /root/reference/src/RecombinationRates.hpp:49
 *   VernerRecombinationRates::get_recombination_rate     /root/reference/src/VernerRecombinationRates.cpp:157-333
 *   VernerRecombinationRates::get_recombination_rate_verner (rrfit)             ...:116-147
 *   ChargeTransferRates::get_charge_transfer_recombination_rate_H               ChargeTransferRates.cpp:44-181
 *   ChargeTransferRates::get_charge_transfer_ionization_rate_H                  ...:191-276
 *   ChargeTransferRates::get_charge_transfer_recombination_rate_He              ...:289-395
 *
 * The fits are the published ones (Verner & Ferland 1996; Nussbaumer & Storey
 * 1983/1987; Mazzotta et al. 1998; Abdel-Naby et al. 2012; Kingdon & Ferland
 * 1996; Arnaud & Rothenflug 1985).  They are table-driven here: one coefficient
 * row per ion instead of one switch arm per ion, so that all lanes of a warp run
 * the same instruction stream regardless of which ion they evaluate.
 */
#pragma once
#include "cmib_common.cuh"
#include "tables.cuh"

namespace cmib {

enum RecombinationKind : int { RR_FIXED = 0, RR_VERNER = 1 };

/* radiative part for the 12 metal ions: index m = ion - ION_C_p1 */
CMIB_HD double rrfit_metal(int m, double T) {
  const double *p = CMIB_TBL(RRFIT)[m];
  if (p[0] != 0.) {
    const double tt = sqrt(T * p[3]);
    return p[1] / (tt * fpow(tt + 1., 1. - p[2]) * fpow(1. + sqrt(T * p[4]), 1. + p[2]));
  }
  return p[1] * fpow(T * 1.e-4, -p[2]);
}

/* Nussbaumer & Storey dielectronic fit coefficients (a, b, c, d, f):
 *   1e-12 * (a/T4 + b + c*T4 + d*T4^2) * T4^-1.5 * exp(-f/T4)
 * rows: C+ C++ N0 N+ N++ O0 O+ Ne+ ; N0 has no a-term and divides f by T4 */
struct DielectronicNS { double a, b, c, d, f; };

CMIB_HD double dielectronic_ns(double a, double b, double c, double d, double f, double T4) {
  const double T4_inv = 1. / T4;
  return 1.e-12 * (a * T4_inv + b + c * T4 + d * T4 * T4) * fpow(T4, -1.5) * exp(-f * T4_inv);
}

CMIB_HD double verner_recombination_rate(int ion, double T) {
  double rate = 0.;
  const double T4 = T * 1.e-4;
  switch (ion) {
  case ION_H_n: {
    const double T1 = T / 3.148;
    const double T2 = T / 7.036e5;
    rate = 7.982e-11 / (sqrt(T1) * fpow(1. + sqrt(T1), 0.252) * fpow(1. + sqrt(T2), 1.748));
    break;
  }
  case ION_He_n: {
    const double T1 = T / 15.54;
    const double T2 = T / 3.676e7;
    rate = 3.294e-11 / (sqrt(T1) * fpow(1. + sqrt(T1), 0.309) * fpow(1. + sqrt(T2), 1.691));
    break;
  }
  case ION_C_p1:
    rate = rrfit_metal(0, T) + dielectronic_ns(1.8267, 4.1012, 4.8443, 0.2261, 0.5960, T4);
    break;
  case ION_C_p2:
    rate = rrfit_metal(1, T) + dielectronic_ns(2.3196, 10.7328, 6.8830, -0.1824, 0.4101, T4);
    break;
  case ION_N_n:
    /* no 1/T4 term, and the exponent is written -0.4398 / T4 in the reference */
    rate = rrfit_metal(2, T) + 1.e-12 * (0.6310 + 0.1990 * T4 - 0.0197 * T4 * T4) *
                                   fpow(T4, -1.5) * exp(-0.4398 / T4);
    break;
  case ION_N_p1:
    rate = rrfit_metal(3, T) + dielectronic_ns(0.0320, -0.6624, 4.3191, 0.0003, 0.5946, T4);
    break;
  case ION_N_p2:
    rate = rrfit_metal(4, T) + dielectronic_ns(-0.8806, 11.2406, 30.7066, -1.1721, 0.6127, T4);
    break;
  case ION_O_n:
    rate = rrfit_metal(5, T) + dielectronic_ns(-0.0001, 0.0001, 0.0956, 0.0193, 0.4106, T4);
    break;
  case ION_O_p1:
    rate = rrfit_metal(6, T) + dielectronic_ns(-0.0036, 0.7519, 1.5252, -0.0838, 0.2769, T4);
    break;
  case ION_Ne_n:
    rate = rrfit_metal(7, T);
    break;
  case ION_Ne_p1:
    rate = rrfit_metal(8, T) + dielectronic_ns(0.0129, -0.1779, 0.9353, -0.0682, 0.4156, T4);
    break;
  case ION_S_p1: {
    const double TeV = T / 1.16045221e4;
    rate = rrfit_metal(9, T) + 1.37e-9 * exp(-14.95 / TeV) * fpow(TeV, -1.5);
    break;
  }
  case ION_S_p2: {
    const double TeV = T / 1.16045221e4;
    const double TeV_inv = 1. / TeV;
    rate = rrfit_metal(10, T) +
           (8.0729e-9 * exp(-17.56 * TeV_inv) + 1.1012e-10 * exp(-7.07 * TeV_inv)) * fpow(TeV, -1.5);
    break;
  }
  case ION_S_p3: {
    const double T_inv = 1. / T;
    rate = rrfit_metal(11, T) +
           (5.817e-7 * exp(-362.8 * T_inv) + 1.391e-6 * exp(-1058. * T_inv) +
            1.123e-5 * exp(-7160. * T_inv) + 1.521e-4 * exp(-3.26e4 * T_inv) +
            1.875e-3 * exp(-1.235e5 * T_inv) + 2.097e-2 * exp(-2.07e5 * T_inv)) *
               fpow(T, -1.5);
    break;
  }
  default:
    break;
  }
  rate *= 1.e-6; /* cm^3 s^-1 -> m^3 s^-1 */
  return rate > 0. ? rate : 0.;
}

/* run-time selected recombination model */
struct RecombinationModel {
  int kind;
  double fixed[NUM_IONS];
};

CMIB_HD double recombination_rate(const RecombinationModel &m, int ion, double T) {
  return m.kind == RR_VERNER ? verner_recombination_rate(ion, T) : m.fixed[ion];
}

/* ---- charge transfer: a * t^b * (1 + c*exp(d*t)) with t clamped to [lo,hi] ---- */
CMIB_HD double ct_clamp(double t, double lo, double hi) {
  double s = (t > lo) ? t : lo; /* std::max(t, lo) */
  s = (s < hi) ? s : hi;        /* std::min(s, hi) */
  return s;
}
CMIB_HD double ct_fit(double a, double b, double c, double d, double t) {
  return a * fpow(t, b) * (1. + c * exp(d * t));
}

/* recombination X^(i+1) + H0 -> X^i + H+ ; T4 = T / 1e4 K */
CMIB_HD double ct_recombination_H(int ion, double T4) {
  switch (ion) {
  case ION_He_n: return ct_fit(7.47e-21, 2.06, 9.93, -3.89, ct_clamp(T4, 0.6, 10.));
  case ION_C_p1: return ct_fit(1.67e-19, 2.79, 304.74, -4.07, ct_clamp(T4, 0.5, 5.));
  case ION_C_p2: return ct_fit(3.25e-15, 0.21, 0.19, -3.29, ct_clamp(T4, 0.1, 10.));
  case ION_N_n: return ct_fit(1.01e-18, -0.29, -0.92, -8.38, ct_clamp(T4, 0.01, 5.));
  case ION_N_p1: return ct_fit(3.05e-16, 0.6, 2.65, -0.93, ct_clamp(T4, 0.1, 10.));
  case ION_N_p2: return ct_fit(4.54e-15, 0.57, -0.65, -0.89, ct_clamp(T4, 0.001, 10.));
  case ION_O_n: return ct_fit(1.04e-15, 3.15e-2, -0.61, -9.73, ct_clamp(T4, 0.001, 1.));
  case ION_O_p1: return ct_fit(1.04e-15, 0.27, 2.02, -5.92, ct_clamp(T4, 0.01, 10.));
  case ION_Ne_n: return 0.;
  case ION_Ne_p1: return 1.e-20;
  case ION_S_p1: return 1.e-20;
  case ION_S_p2: return ct_fit(2.29e-15, 4.02e-2, 1.59, -6.06, ct_clamp(T4, 0.1, 3.));
  case ION_S_p3: return ct_fit(6.44e-15, 0.13, 2.69, -5.69, ct_clamp(T4, 0.1, 3.));
  default: return 0.;
  }
}

/* ionization X^i + H+ -> X^(i+1) + H0 (only N0 and O0 are non-zero) */
CMIB_HD double ct_ionization_H(int ion, double T4) {
  switch (ion) {
  case ION_N_n: {
    const double t = ct_clamp(T4, 0.01, 5.);
    return 4.55e-18 * fpow(t, -0.29) * (1. - 0.92 * exp(-8.38 * t)) * exp(-1.086 / t);
  }
  case ION_O_n: {
    const double t = ct_clamp(T4, 0.001, 1.);
    return 7.4e-17 * fpow(t, 0.47) * (1. + 24.37 * exp(-0.74 * t)) * exp(-0.023 / t);
  }
  default: return 0.;
  }
}

/* recombination X^(i+1) + He0 -> X^i + He+ */
CMIB_HD double ct_recombination_He(int ion, double T4) {
  switch (ion) {
  case ION_C_p2: {
    const double t = ct_clamp(T4, 0.1, 3.);
    return 4.6e-17 * t * t;
  }
  case ION_N_p1: return ct_fit(3.3e-16, 0.29, 1.3, -4.5, ct_clamp(T4, 0.1, 3.));
  case ION_N_p2: return 1.5e-16;
  case ION_O_p1: return 2.e-16 * fpow(ct_clamp(T4, 0.5, 5.), 0.95);
  case ION_Ne_p1: return 1.e-20;
  case ION_S_p2: return 1.1e-15 * fpow(ct_clamp(T4, 0.1, 3.), 0.56);
  case ION_S_p3: return ct_fit(7.6e-19, 0.32, 3.4, -5.25, ct_clamp(T4, 0.1, 3.));
  default: return 0.;
  }
}

} // namespace cmib
