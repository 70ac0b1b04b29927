/*
 * state.cuh — per-cell ionization-state and temperature solve.
 *
 * The closed forms below (H/He balance, the metal ladders, the heating/cooling balance) are a TRANSLITERATION of the
 * reference's arithmetic, operation for operation and in its order: the results have to be bit-equal on the
 * reference's random stream (tests/test_host_physics.py: a whole simulation, all cells, all 14 fractions), which
 * leaves no freedom in the expressions.  What is this repository's own is around them: the resumable per-cell state
 * machine (TemperatureSolve), three lanes per cell, the warp-wide line cooling of a warp's last cell (kernels.cuh).
 *
 * Reference code restated here:
 *   IonizationStateCalculator::calculate_ionization_state(jfac,hfac,cell)
 *                                   /root/reference/src/IonizationStateCalculator.cpp:70-272
 *   ::compute_ionization_states_hydrogen_helium                     ...:649-753
 *   ::compute_ionization_state_hydrogen                             ...:802-820
 *   ::compute_ionization_states_metals                              ...:323-501
 *   TemperatureCalculator::compute_cooling_and_heating_balance
 *                                   /root/reference/src/TemperatureCalculator.cpp:207-501
 *   TemperatureCalculator::calculate_temperature(cell,...)          ...:567-931
 *   grid-level dispatch (do_T && loop > min_iter)                   ...:944-970
 *
 * All quirks of the
 * reference that affect the result are kept: the
 * neutral branch of the ionization-only path sets N0/O0/Ne0 fractions to 1 while
 * the temperature path sets them to 0, the metals written by the LAST balance
 * evaluation (at the pre-update T0) are the ones that survive, etc.
 */
#pragma once
#include "cmib_common.cuh"
#include "linecooling.cuh"
#include "rates.cuh"

namespace cmib {

struct TemperatureParams {
  int do_temperature;          /* TemperatureCalculator:do temperature calculation */
  uint32_t min_iterations;     /* :minimum number of iterations (T solve when loop > this) */
  double epsilon;              /* :epsilon convergence */
  uint32_t max_iterations;     /* :maximum number of iterations */
  double pahfac;               /* :PAH heating factor */
  double crfac;                /* :cosmic ray heating factor */
  double crlim;                /* :cosmic ray heating limit */
  double crscale;              /* :cosmic ray heating scale length (m) */
  double min_ionized_T;        /* :minimum ionized temperature (K) */
};

/* result of one cell update: T, 14 fractions, 2 normalised heating terms */
struct CellState {
  double T;
  double x[NUM_IONS];
  double heat[NUM_HEAT];
};

CMIB_HD double ionization_state_hydrogen(double alphaH, double jH, double nH) {
  if (jH > 0. && nH > 0.) {
    const double aa = 0.5 * jH / (nH * alphaH);
    const double bb = 2. / aa;
    if (bb < 1.e-10) {
      const double v = 0.25 * bb;
      return (1.e-14 < v) ? v : 1.e-14; /* std::max(1e-14, v) */
    }
    const double cc = sqrt(bb + 1.);
    const double v = 1. + aa * (1. - cc);
    return (1.e-14 < v) ? v : 1.e-14;
  }
  return 1.;
}

/* coupled H/He fixed point; returns number of iterations (>20 flags the
 * reference's "Too many iterations" fatal error, here reported not aborted) */
CMIB_HD int ionization_states_hydrogen_helium(double alphaH, double alphaHe, double jH, double jHe,
                                              double nH, double AHe, double T, double &h0,
                                              double &he0) {
  if (jH < 1.e-20) {
    h0 = 1.;
    he0 = 1.;
    return 0;
  }
  const double alpha_e_2sP = 4.17e-20 * fpow(T * 1.e-4, -0.861);
  const double ch1 = alphaH * nH / jH;
  const double ch2 = AHe * alpha_e_2sP * nH / jH;
  double che = 0.;
  if (jHe > 0.) che = alphaHe * nH / jHe;
  double h0old = 0.99 * (1. - exp(-0.5 / ch1));
  h0 = 0.9 * h0old;
  double he0old = 1.;
  if (che > 0.) {
    he0old = 0.5 / che;
    he0old = (1. < he0old) ? 1. : he0old; /* std::min(he0old, 1.) */
  }
  he0 = 0.;
  int niter = 0;
  const double sqrtT = sqrt(T);
  while (fabs(h0 - h0old) > 1.e-4 * h0old && fabs(he0 - he0old) > 1.e-4 * he0old) {
    ++niter;
    h0old = h0;
    he0old = (he0 > 0.) ? he0 : 0.;
    const double pHots = 1. / (1. + 77. * he0old / sqrtT / h0old);
    const double ch = ch1 - ch2 * AHe * (1. - he0old) * pHots / (1. - h0old);
    he0 = 1.;
    if (che != 0.) {
      const double bhe = (1. + 2. * AHe - h0) * che + 1.;
      const double che_bhe = che / bhe;
      const double opAHeh0 = 1. + AHe - h0;
      const double t1he = 4. * AHe * opAHeh0 * che_bhe * che_bhe;
      if (t1he < 1.e-3) {
        he0 = opAHeh0 * che_bhe;
      } else {
        he0 = (bhe - sqrt(bhe * bhe - 4. * AHe * opAHeh0 * che * che)) / (2. * AHe * che);
      }
    }
    const double b = ch * (2. + AHe - he0 * AHe) + 1.;
    const double ch_b = ch / b;
    const double opAHeh0AHe = 1. + AHe - he0 * AHe;
    const double t1 = 4. * ch_b * ch_b * opAHeh0AHe;
    if (t1 < 1.e-3) {
      h0 = ch_b * opAHeh0AHe;
    } else {
      h0 = (b - sqrt(b * b - 4. * ch * ch * opAHeh0AHe)) / (2. * ch);
    }
    if (niter > 10) {
      h0 = 0.5 * (h0 + h0old);
      he0 = 0.5 * (he0 + he0old);
    }
    if (niter > 20) break;
  }
  return niter;
}

/* metals ladder balance; jm = 12 normalised mean intensities C+ .. S+++ ;
 * writes x[ION_C_p1 .. ION_S_p3] */
CMIB_HD void ionization_states_metals(const double *jm, double ne, double T, double T4,
                                      double nh0, double nhe0, double nhp,
                                      const RecombinationModel &rr, double *x) {
  {
    const double a0 = recombination_rate(rr, ION_C_p1, T);
    const double a1 = recombination_rate(rr, ION_C_p2, T);
    const double C21 = jm[0] / (ne * a0);
    const double C32 = jm[1] / (ne * a1 + nh0 * ct_recombination_H(ION_C_p2, T4) +
                                nhe0 * ct_recombination_He(ION_C_p2, T4));
    const double C31 = C32 * C21;
    const double s = 1. / (1. + C21 + C31);
    x[ION_C_p1] = C21 * s;
    x[ION_C_p2] = C31 * s;
  }
  {
    const double a0 = recombination_rate(rr, ION_N_n, T);
    const double a1 = recombination_rate(rr, ION_N_p1, T);
    const double a2 = recombination_rate(rr, ION_N_p2, T);
    const double N21 = (jm[2] + nhp * ct_ionization_H(ION_N_n, T4)) /
                       (ne * a0 + nh0 * ct_recombination_H(ION_N_n, T4));
    const double N32 = jm[3] / (ne * a1 + nh0 * ct_recombination_H(ION_N_p1, T4) +
                                nhe0 * ct_recombination_He(ION_N_p1, T4));
    const double N43 = jm[4] / (ne * a2 + nh0 * ct_recombination_H(ION_N_p2, T4) +
                                nhe0 * ct_recombination_He(ION_N_p2, T4));
    const double N31 = N32 * N21;
    const double N41 = N43 * N31;
    const double s = 1. / (1. + N21 + N31 + N41);
    x[ION_N_n] = N21 * s;
    x[ION_N_p1] = N31 * s;
    x[ION_N_p2] = N41 * s;
  }
  {
    const double a0 = recombination_rate(rr, ION_O_n, T);
    const double a1 = recombination_rate(rr, ION_O_p1, T);
    const double O21 = (jm[5] + nhp * ct_ionization_H(ION_O_n, T4)) /
                       (ne * a0 + nh0 * ct_recombination_H(ION_O_n, T4));
    const double O32 = jm[6] / (ne * a1 + nh0 * ct_recombination_H(ION_O_p1, T4) +
                                nhe0 * ct_recombination_He(ION_O_p1, T4));
    const double O31 = O32 * O21;
    const double s = 1. / (1. + O21 + O31);
    x[ION_O_n] = O21 * s;
    x[ION_O_p1] = O31 * s;
  }
  {
    const double a0 = recombination_rate(rr, ION_Ne_n, T);
    const double a1 = recombination_rate(rr, ION_Ne_p1, T);
    const double Ne21 = jm[7] / (ne * a0);
    const double Ne32 = jm[8] / (ne * a1 + nh0 * ct_recombination_H(ION_Ne_p1, T4) +
                                 nhe0 * ct_recombination_He(ION_Ne_p1, T4));
    const double Ne31 = Ne32 * Ne21;
    const double s = 1. / (1. + Ne21 + Ne31);
    x[ION_Ne_n] = Ne21 * s;
    x[ION_Ne_p1] = Ne31 * s;
  }
  {
    const double a0 = recombination_rate(rr, ION_S_p1, T);
    const double a1 = recombination_rate(rr, ION_S_p2, T);
    const double a2 = recombination_rate(rr, ION_S_p3, T);
    const double S21 = jm[9] / (ne * a0 + nh0 * ct_recombination_H(ION_S_p1, T4));
    const double S32 = jm[10] / (ne * a1 + nh0 * ct_recombination_H(ION_S_p2, T4) +
                                 nhe0 * ct_recombination_He(ION_S_p2, T4));
    const double S43 = jm[11] / (ne * a2 + nh0 * ct_recombination_H(ION_S_p3, T4) +
                                 nhe0 * ct_recombination_He(ION_S_p3, T4));
    const double S31 = S32 * S21;
    const double S41 = S43 * S31;
    const double s = 1. / (1. + S21 + S31 + S41);
    x[ION_S_p1] = S21 * s;
    x[ION_S_p2] = S31 * s;
    x[ION_S_p3] = S41 * s;
  }
}

/* ionization-only update (temperature fixed).  J / heat are the raw
 * accumulators; jfac/hfac already divided by the cell volume. */
CMIB_HD void cell_ionization_state(double jfac, double hfac, const double *J, const double *heat,
                                   double ntot, double T, const double *abund,
                                   const RecombinationModel &rr, CellState &out) {
  const double jH = jfac * J[ION_H_n];
  const double jHe = jfac * J[ION_He_n];
  out.heat[HEAT_H] = hfac * heat[HEAT_H];
  out.heat[HEAT_He] = hfac * heat[HEAT_He];
  out.T = T;
  if (jH > 0. && ntot > 0.) {
    const double alphaH = recombination_rate(rr, ION_H_n, T);
    const double AHe = abund[EL_He];
    double h0, he0 = 0.;
    if (AHe != 0.) {
      const double alphaHe = recombination_rate(rr, ION_He_n, T);
      ionization_states_hydrogen_helium(alphaH, alphaHe, jH, jHe, ntot, AHe, T, h0, he0);
    } else {
      h0 = ionization_state_hydrogen(alphaH, jH, ntot);
    }
    out.x[ION_H_n] = h0;
    out.x[ION_He_n] = he0;
    const double nhp = ntot * (1. - h0);
    const double ne = ntot * (1. - h0 + AHe * (1. - he0));
    const double T4 = T * 1.e-4;
    double jm[12];
#pragma unroll
    for (int m = 0; m < 12; ++m) jm[m] = jfac * J[2 + m];
    const double nh0 = ntot * h0;
    const double nhe0 = ntot * he0 * AHe;
    ionization_states_metals(jm, ne, T, T4, nh0, nhe0, nhp, rr, out.x);
  } else if (ntot > 0.) {
    /* neutral cell: note N0, O0, Ne0 = 1 here (IonizationStateCalculator.cpp:190-224) */
#pragma unroll
    for (int i = 0; i < NUM_IONS; ++i) out.x[i] = 0.;
    out.x[ION_H_n] = 1.;
    out.x[ION_He_n] = 1.;
    out.x[ION_N_n] = 1.;
    out.x[ION_O_n] = 1.;
    out.x[ION_Ne_n] = 1.;
  } else {
#pragma unroll
    for (int i = 0; i < NUM_IONS; ++i) out.x[i] = 0.;
  }
}

/* heating/cooling balance at temperature T.  j[14], h[2] normalised.  Writes the
 * 12 metal fractions into xm (index by Ion). */
/* what the balance hands from its first half (ionization states, heating) over the line cooling to its second half
 * (the remaining cooling terms): the kernel evaluates the line cooling of its last cells warp-wide in between */
struct BalanceMid {
  double T, n, ne, nenhp, nenhep, sqrtT, logT;
  double ab[LC_NUM];
};

CMIB_HD void balance_before_line_cooling(double &h0, double &he0, double &gain, BalanceMid &mid, double T, double n,
                                         double midz, const double *j, const double *abund, const double *h,
                                         double pahfac, double crfac, double crscale, const RecombinationModel &rr,
                                         double *xm) {
  const double alphaH = recombination_rate(rr, ION_H_n, T);
  const double alphaHe = recombination_rate(rr, ION_He_n, T);
  const double jH = j[ION_H_n];
  const double jHe = j[ION_He_n];
  const double hH = h[HEAT_H];
  const double hHe = h[HEAT_He];
  const double T4 = T * 1.e-4;
  const double sqrtT = sqrt(T);
  const double logT = log(T);
  const double AHe = abund[EL_He];

  ionization_states_hydrogen_helium(alphaH, alphaHe, jH, jHe, n, AHe, T, h0, he0);

  const double ne = n * (1. - h0 + AHe * (1. - he0));
  const double nhp = n * (1. - h0);
  const double nhep = (1. - he0) * n * AHe;
  const double nenhp = ne * nhp;
  const double nenhep = ne * nhep;

  gain = n * (hH * h0 + hHe * AHe * he0);
  const double alpha_e_2sP = 4.17e-20 * fpow(T4, -0.861);
  const double pHots = 1. / (1. + 77. * he0 / (sqrtT * h0));
  gain += pHots * 1.21765423e-18 * alpha_e_2sP * nenhep;
  gain += 1.5e-37 * n * ne * pahfac;
  double heatcr = 0.;
  if (crfac > 0.) {
    heatcr = crfac * 1.2e-25 / sqrt(ne);
    if (crscale > 0.) heatcr *= exp(-fabs(midz) / crscale);
  }
  gain += heatcr;

  const double nh0 = n * h0;
  const double nhe0 = n * he0 * AHe;
  ionization_states_metals(j + 2, ne, T, T4, nh0, nhe0, nhp, rr, xm);

  double *ab = mid.ab;
  ab[LC_CII] = abund[EL_C] * (1. - xm[ION_C_p1] - xm[ION_C_p2]);
  ab[LC_CIII] = abund[EL_C] * xm[ION_C_p1];
  ab[LC_NI] = abund[EL_N] * (1. - xm[ION_N_n] - xm[ION_N_p1] - xm[ION_N_p2]);
  ab[LC_NII] = abund[EL_N] * xm[ION_N_n];
  ab[LC_NIII] = abund[EL_N] * xm[ION_N_p1];
  ab[LC_OI] = abund[EL_O] * (1. - xm[ION_O_n] - xm[ION_O_p1]);
  ab[LC_OII] = abund[EL_O] * xm[ION_O_n];
  ab[LC_OIII] = abund[EL_O] * xm[ION_O_p1];
  ab[LC_NeII] = abund[EL_Ne] * xm[ION_Ne_n];
  ab[LC_NeIII] = abund[EL_Ne] * xm[ION_Ne_p1];
  ab[LC_SII] = abund[EL_S] * (1. - xm[ION_S_p1] - xm[ION_S_p2] - xm[ION_S_p3]);
  ab[LC_SIII] = abund[EL_S] * xm[ION_S_p1];
  ab[LC_SIV] = abund[EL_S] * xm[ION_S_p2];
  mid.T = T; mid.n = n; mid.ne = ne; mid.nenhp = nenhp; mid.nenhep = nenhep; mid.sqrtT = sqrtT; mid.logT = logT;
}

/* cooling = the line cooling per hydrogen atom (line_cooling(mid.T, mid.ne, mid.ab)) */
CMIB_HD void balance_after_line_cooling(double &gain, double &loss, double cooling, const BalanceMid &mid) {
  const double T = mid.T, sqrtT = mid.sqrtT, logT = mid.logT, nenhp = mid.nenhp, nenhep = mid.nenhep;
  loss = cooling * mid.n;

  const double c = 5.5 - logT;
  const double gff = 1.1 + 0.34 * exp(-c * c / 3.);
  loss += 1.42e-40 * gff * sqrtT * (nenhp + nenhep);
  const double Lhp = 2.85e-40 * nenhp * sqrtT * (5.914 - 0.5 * logT + 0.01184 * cbrt(T));
  const double Lhep = 1.55e-39 * nenhep * powl(T, logT, 0.3647);
  loss += Lhp + Lhep;
  loss = (loss < 0.) ? 0. : loss; /* std::max(loss, 0.) */
  gain = (gain < 0.) ? 0. : gain;
}

CMIB_HD void cooling_heating_balance(double &h0, double &he0, double &gain, double &loss, double T,
                                     double n, double midz, const double *j, const double *abund,
                                     const double *h, double pahfac, double crfac, double crscale,
                                     const RecombinationModel &rr, double *xm) {
  BalanceMid mid;
  balance_before_line_cooling(h0, he0, gain, mid, T, n, midz, j, abund, h, pahfac, crfac, crscale, rr, xm);
  balance_after_line_cooling(gain, loss, line_cooling(mid.T, mid.ne, mid.ab), mid);
}

CMIB_HD void set_neutral_T(CellState &out) {
  out.T = 500.;
#pragma unroll
  for (int i = 0; i < NUM_IONS; ++i) out.x[i] = 0.;
  out.x[ION_H_n] = 1.;
  out.x[ION_He_n] = 1.;
  out.heat[HEAT_H] = 0.;
  out.heat[HEAT_He] = 0.;
}

/*
 * Temperature solve of one cell as a resumable state machine: every step is ONE evaluation of
 * the heating/cooling balance (the expensive part) at the temperature temperature_solve_T()
 * names, followed by temperature_solve_advance().  TemperatureCalculator::calculate_temperature
 * (TemperatureCalculator.cpp:567-931) evaluates the balance at 1.1 T0, 0.9 T0 and T0 per secant
 * iteration; cells need 1 to 100 iterations.  The kernel keeps all lanes of a warp inside the
 * balance evaluation and hands a new cell to a lane as soon as its cell has converged
 * (update_temperature_kernel), instead of waiting for the slowest cell of the warp.
 * cell_temperature() below runs the same pieces in a plain loop (host logic check, eval probes).
 */
struct TemperatureSolve {
  double T0, crfac;
  double gain0, loss0, gain1, loss1, gain2, loss2, h0, he0;
  uint32_t niter;
  int phase; /* 0: balance at 1.1 T0, 1: at 0.9 T0, 2: at T0, then the secant update */
};

/* returns false when the cell needs no solve (out is final); otherwise fills j[14], h[2] with the
 * normalised mean intensities / heating terms, seeds out.x with the cell's current fractions */
CMIB_HD bool temperature_solve_begin(TemperatureSolve &S, double jfac, double hfac, const double *J,
                                     const double *heat, double ntot, double Tcell, double cr_factor,
                                     const double *abund, const RecombinationModel &rr,
                                     const TemperatureParams &tp, const double *xprev, double *j,
                                     double *h, CellState &out) {
  const double jH = jfac * J[ION_H_n];
  const double jHe = jfac * J[ION_He_n];
  if ((jH == 0. && jHe == 0.) || ntot == 0.) {
    set_neutral_T(out);
    return false;
  }
  double crfac = tp.crfac * cr_factor;
  if (crfac < 0.) crfac = tp.crfac;
  if (crfac > 0.) {
    double h0, he0;
    const double alphaH = recombination_rate(rr, ION_H_n, 8000.);
    const double alphaHe = recombination_rate(rr, ION_He_n, 8000.);
    ionization_states_hydrogen_helium(alphaH, alphaHe, jH, jHe, ntot, abund[EL_He], 8000., h0, he0);
    if (h0 > tp.crlim) {
      set_neutral_T(out);
      return false;
    }
  }
  S.crfac = crfac;
  S.T0 = Tcell;
  if (Tcell <= 4000.) S.T0 = 8000.;
#pragma unroll
  for (int i = 0; i < NUM_IONS; ++i) j[i] = jfac * J[i];
  h[HEAT_H] = hfac * heat[HEAT_H];
  h[HEAT_He] = hfac * heat[HEAT_He];
#pragma unroll
  for (int i = 0; i < NUM_IONS; ++i) out.x[i] = xprev[i];
  S.niter = 0;
  S.gain0 = 1.;
  S.loss0 = 0.;
  S.h0 = 0.;
  S.he0 = 0.;
  S.gain1 = S.loss1 = S.gain2 = S.loss2 = 0.;
  S.phase = 0;
  return true;
}

/* the loop condition of the reference (:730): true while another secant iteration is due */
CMIB_HD bool temperature_solve_continues(const TemperatureSolve &S, const TemperatureParams &tp) {
  return fabs(S.gain0 - S.loss0) > tp.epsilon * S.gain0 && S.niter < tp.max_iterations;
}

CMIB_HD double temperature_solve_T(const TemperatureSolve &S) {
  return (S.phase == 0) ? 1.1 * S.T0 : ((S.phase == 1) ? 0.9 * S.T0 : S.T0);
}

/* the secant update of one iteration from its three balances (1: at 1.1 T0, 2: at 0.9 T0, 0: at T0);
 * returns true when the solve has finished */
CMIB_HD bool temperature_solve_update(TemperatureSolve &S, double gain1, double loss1, double gain2,
                                      double loss2, double h0e, double he0e, double gain0, double loss0,
                                      const TemperatureParams &tp) {
  S.gain1 = gain1;
  S.loss1 = loss1;
  S.gain2 = gain2;
  S.loss2 = loss2;
  S.h0 = h0e;
  S.he0 = he0e;
  S.gain0 = gain0;
  S.loss0 = loss0;
  S.phase = 0;
  double expgain;
  if (S.gain2 > 0.) {
    expgain = (S.gain1 > 0.) ? log(S.gain1 / S.gain2) : -99.;
  } else {
    expgain = (S.gain1 > 0.) ? 99. : 0.;
  }
  double exploss;
  if (S.loss2 > 0.) {
    exploss = (S.loss1 > 0.) ? log(S.loss1 / S.loss2) : -99.;
  } else {
    exploss = (S.loss1 > 0.) ? 99. : 0.;
  }
  const double expdiff = expgain - exploss;
  const double logtt = log(1.1 / 0.9);
  if (S.gain0 > 0. && expdiff != 0.) {
    S.T0 *= pow(S.loss0 / S.gain0, logtt / expdiff);
  } else {
    S.T0 = 1.1 * S.T0;
  }
  if (S.T0 < tp.min_ionized_T) {
    S.T0 = 500.;
    S.h0 = 1.;
    S.he0 = 1.;
    S.gain0 = 1.;
    S.loss0 = 1.;
  }
  if (S.T0 > 1.e10) {
    S.T0 = 1.e10;
    S.h0 = 1.e-10;
    S.he0 = 1.e-10;
    S.gain0 = 1.;
    S.loss0 = 1.;
  }
  return !temperature_solve_continues(S, tp);
}

/* record the balance just evaluated; after the third one do the secant update.  Returns true
 * when the solve has finished. */
CMIB_HD bool temperature_solve_advance(TemperatureSolve &S, double h0e, double he0e, double gain,
                                       double loss, const TemperatureParams &tp) {
  if (S.phase == 0) {
    ++S.niter;
    S.gain1 = gain;
    S.loss1 = loss;
    S.phase = 1;
    return false;
  }
  if (S.phase == 1) {
    S.gain2 = gain;
    S.loss2 = loss;
    S.phase = 2;
    return false;
  }
  return temperature_solve_update(S, S.gain1, S.loss1, S.gain2, S.loss2, h0e, he0e, gain, loss, tp);
}

/* h = the normalised heating terms of temperature_solve_begin */
CMIB_HD void temperature_solve_finish(const TemperatureSolve &S, const double *J, const double *h,
                                      CellState &out) {
  double T0 = S.T0, h0 = S.h0, he0 = S.he0;
  T0 = (T0 < 30000.) ? T0 : 30000.; /* std::min(30000., T0) */
  out.T = T0;
  if (J[ION_H_n] == 0.) h0 = 1.;
  if (J[ION_He_n] == 0.) he0 = 1.;
  out.x[ION_H_n] = h0;
  out.x[ION_He_n] = he0;
  if (h0 == 1. || h0 <= 1.e-10) {
#pragma unroll
    for (int i = 2; i < NUM_IONS; ++i) out.x[i] = 0.;
  }
  out.heat[HEAT_H] = h[HEAT_H];
  out.heat[HEAT_He] = h[HEAT_He];
}

/* full temperature + ionization update of one cell; `xprev` = current metal
 * fractions of the cell (kept when no balance evaluation overwrites them). */
CMIB_HD void cell_temperature(double jfac, double hfac, const double *J, const double *heat,
                              double ntot, double Tcell, double cr_factor, double midz,
                              const double *abund, const RecombinationModel &rr,
                              const TemperatureParams &tp, const double *xprev, CellState &out) {
  TemperatureSolve S;
  double j[NUM_IONS], h[NUM_HEAT];
  if (!temperature_solve_begin(S, jfac, hfac, J, heat, ntot, Tcell, cr_factor, abund, rr, tp, xprev, j, h, out))
    return;
  bool finished = !temperature_solve_continues(S, tp);
  while (!finished) {
    double h0e, he0e, gain, loss;
    cooling_heating_balance(h0e, he0e, gain, loss, temperature_solve_T(S), ntot, midz, j, abund, h,
                            tp.pahfac, S.crfac, tp.crscale, rr, out.x);
    finished = temperature_solve_advance(S, h0e, he0e, gain, loss, tp);
  }
  temperature_solve_finish(S, J, h, out);
}

} // namespace cmib
