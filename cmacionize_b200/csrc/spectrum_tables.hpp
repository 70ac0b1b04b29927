/*
 * spectrum_tables.hpp — host-side construction of the tabulated spectra that the
 * device samples from.  Built once per configuration on the host (they are
 * O(1e5) doubles and depend on libm's exp/log10, so building them on the host
 * with the reference's operation order makes them bit-identical to the
 * reference's tables), then uploaded.
 *
 * Behavioural contract:
 *   PlanckPhotonSourceSpectrum ctor        /root/reference/src/PlanckPhotonSourceSpectrum.cpp:53-115
 *   HydrogenLymanContinuumSpectrum ctor    /root/reference/src/HydrogenLymanContinuumSpectrum.cpp:40-122
 *   HeliumLymanContinuumSpectrum ctor      /root/reference/src/HeliumLymanContinuumSpectrum.cpp:40-133
 *   HeliumTwoPhotonContinuumSpectrum ctor  /root/reference/src/HeliumTwoPhotonContinuumSpectrum.cpp:44-101
 * including the quirks listed in SURVEY.md Appendix C (two different "13.6 eV in
 * Hz" literals; lower-edge emissivity divided by upper-edge frequency).
 */
#pragma once
#include <cmath>
#include <functional>
#include <vector>

#include "cmib_common.cuh"
#include "source.cuh"

namespace cmib {
namespace host {

/* out: [3][1000] = cdf, log10(cdf), log10(nu / 13.6 eV) */
inline void build_planck_table(double temperature, std::vector<double> &out) {
  const int N = SPECTRUM_NUMFREQ;
  out.assign(3 * N, 0.);
  double *cdf = out.data(), *logcdf = cdf + N, *lognu = cdf + 2 * N;
  const double nu_unit = 3.289e15; /* the table-build literal, not the sampling one */
  std::vector<double> x(N), lum(N);
  for (int i = 0; i < N; ++i) {
    x[i] = 1. + i * (4. - 1.) / (N - 1.);
    lum[i] = x[i] * x[i] * x[i] / (std::exp(PLANCK * x[i] * nu_unit / (BOLTZMANN * temperature)) - 1.);
  }
  for (int i = 1; i < N; ++i)
    cdf[i] = cdf[i - 1] + 0.5 * (lum[i] / x[i] + lum[i - 1] / x[i - 1]) * (x[i] - x[i - 1]);
  logcdf[0] = -10.;
  lognu[0] = 0.;
  for (int i = 1; i < N; ++i) {
    cdf[i] /= cdf[N - 1];
    logcdf[i] = std::log10(cdf[i]);
    lognu[i] = std::log10(x[i]);
  }
}

/* Lyman-continuum re-emission tables for H (ion 0) or He (ion 1).
 * sigma(nu) is the photoionization cross section of that ion. */
inline void build_lyc_table(int which, const std::function<double(double)> &sigma,
                            std::vector<double> &freq, std::vector<double> &temp,
                            std::vector<double> &cdf) {
  const int N = SPECTRUM_NUMFREQ, NT = LYC_NUMTEMP;
  const double nu_min = (which == 0) ? 3.289e15 : 1.81 * 3.288465385e15;
  const double nu_max = (which == 0) ? 4. * nu_min : 4. * 3.288465385e15;
  freq.assign(N, 0.);
  temp.assign(NT, 0.);
  cdf.assign((size_t)NT * N, 0.);
  for (int i = 0; i < N; ++i) freq[i] = nu_min + i * (nu_max - nu_min) / (N - 1.);
  /* the cross section does not depend on T: evaluate it once per frequency */
  std::vector<double> xs(N);
  for (int i = 0; i < N; ++i) xs[i] = sigma(freq[i]);
  for (int iT = 0; iT < NT; ++iT) {
    double *row = cdf.data() + (size_t)iT * N;
    temp[iT] = 1500. + (iT + 0.5) * 13500. / NT;
    for (int i = 1; i < N; ++i) {
      const double lo = freq[i - 1], hi = freq[i];
      const double jlo = lo * lo * lo * xs[i - 1] *
                         std::exp(-(PLANCK * (lo - nu_min)) / (BOLTZMANN * temp[iT]));
      const double jhi = hi * hi * hi * xs[i] *
                         std::exp(-(PLANCK * (hi - nu_min)) / (BOLTZMANN * temp[iT]));
      /* sic: lower-edge emissivity over upper-edge frequency and vice versa */
      row[i] = 0.5 * (jlo / hi + jhi / lo) * (hi - lo);
    }
    for (int i = 1; i < N; ++i) row[i] = row[i - 1] + row[i];
    for (int i = 0; i < N; ++i) row[i] /= row[N - 1];
  }
}

inline double he2q_interpolate(double y) {
  if (!(y < 1.)) return 0.;
  const uint32_t k = locate(y, h_HE2Q_Y, 41);
  const double f = (y - h_HE2Q_Y[k]) / (h_HE2Q_Y[k + 1] - h_HE2Q_Y[k]);
  return h_HE2Q_A[k] + f * (h_HE2Q_A[k + 1] - h_HE2Q_A[k]);
}

inline void build_he2pc_table(std::vector<double> &freq, std::vector<double> &cdf) {
  const int N = SPECTRUM_NUMFREQ;
  freq.assign(N, 0.);
  cdf.assign(N, 0.);
  const double nu_min = 3.288465385e15;
  const double nu_max = 1.6 * nu_min;
  const double nu0 = 4.98e15;
  for (int i = 0; i < N; ++i) freq[i] = nu_min + i * (nu_max - nu_min) / (N - 1.);
  for (int i = 1; i < N; ++i) {
    const double a1 = he2q_interpolate(freq[i - 1] / nu0);
    const double a2 = he2q_interpolate(freq[i] / nu0);
    cdf[i] = 0.5 * (a1 + a2) * (freq[i] - freq[i - 1]);
  }
  for (int i = 1; i < N; ++i) cdf[i] = cdf[i - 1] + cdf[i];
  for (int i = 0; i < N; ++i) cdf[i] /= cdf[N - 1];
}

/* bracket guide of a non-decreasing CDF row for locate_guided (source.cuh):
 * guide[g] = largest j with cdf[j] < g / G, 0 if there is none; g = 0 .. G */
inline void build_guide(const double *cdf, int n, int G, uint16_t *guide) {
  int j = 0;
  for (int g = 0; g <= G; ++g) {
    const double edge = (double)g / (double)G;
    while (j + 1 < n && cdf[j + 1] < edge) ++j;
    guide[g] = (uint16_t)((cdf[j] < edge) ? j : 0);
  }
}

} // namespace host
} // namespace cmib
