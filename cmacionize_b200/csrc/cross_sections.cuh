/*
 * cross_sections.cuh — photoionization cross sections sigma_ion(nu) for the 14
 * tracked ions.
 *
 * Behavioural contract:
 *   CrossSections::get_cross_section(ion, nu)   /root/reference/src/CrossSections.hpp:49
 *   VernerCrossSections::get_cross_section      /root/reference/src/VernerCrossSections.cpp:259-322
 *   VernerCrossSections::get_cross_section_verner                      ...:166-245
 *   FixedValueCrossSections                     /root/reference/src/FixedValueCrossSections.hpp
 *
 * Design: the reference walks three nested std::vector tables per call; here the
 * 23 (Z,N,shell) records that the 14 ions can ever touch are flattened at build
 * time (tools/gen_atomic_data.py) into one [14][2][18] constant table with the
 * shell-selection logic (nout / nint / E_inn) already resolved, so a lane does
 * one uniform constant-cache read per shell and at most two pow() per shell it
 * is actually above threshold for.
 */
#pragma once
#include "cmib_common.cuh"
#include "tables.cuh"

namespace cmib {

enum CrossSectionKind : int { XS_FIXED = 0, XS_VERNER = 1, XS_BIMODAL = 2 };

/* A * x^a * z^b.  Host build: the reference's expression with two pow(), left to right, so that
 * the table transcription and the re-emission spectra tabulated from it are pinned bit for bit
 * (tests/test_host_physics.py).  Device: A * exp(a ln x + b ln z) — one exp and two logs instead of
 * two pows (pow is ~2.5x a log+exp pair in FP64 on sm_100a and the 46 pows of a packet's 14 cross
 * sections were 2/3 of the emission kernel, profiles/r01_prepare.md).  The exponent is < ~60 in
 * magnitude, so the result moves by <= ~2e-14 relative, five orders of magnitude inside the 1e-9
 * tolerance of the reference's own golden test (test/testVernerCrossSections.cpp) and measured at
 * every GPU-tier run. */
CMIB_HD double shell_profile(double A, double x, double a, double z, double b) {
#if defined(__CUDA_ARCH__)
  return A * exp(a * log(x) + b * log(z));
#else
  return A * pow(x, a) * pow(z, b);
#endif
}

/* one shell of phfit2: returns the partial cross section (m^2) at frequency e (Hz) */
CMIB_HD double verner_shell(const double *r, double e) {
  /* r layout: see tools/gen_atomic_data.py */
  if (r[0] == 0.) return 0.;
  if (e < r[3]) return 0.;          /* below the shell threshold */
  const double einn = r[4];
  if (r[1] != 0. && e < einn) return 0.;
  if (r[2] != 0. || e >= einn) {
    const double y = e * r[5];
    const double ym1 = y - 1.;
    const double Fy = shell_profile(ym1 * ym1 + r[9], y, r[10], 1. + sqrt(y * r[7]), -r[8]);
    return r[6] * Fy;
  } else {
    const double x = e * r[11] - r[16];
    const double y = sqrt(x * x + r[17]);
    const double xm1 = x - 1.;
    const double P = r[14];
    const double Fy = shell_profile(xm1 * xm1 + r[15], y, 0.5 * P - 5.5, 1. + sqrt(y * r[13]), -P);
    return r[12] * Fy;
  }
}

CMIB_HD double verner_cross_section(int ion, double e) {
  const double(*tab)[2][18] = CMIB_TBL(VERNER_SHELLS);
  const double a = verner_shell(tab[ion][0], e);
  if (tab[ion][1][0] == 0.) return a;
  return a + verner_shell(tab[ion][1], e);
}

} // namespace cmib
