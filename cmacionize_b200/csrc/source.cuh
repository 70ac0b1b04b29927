/*
 * source.cuh — photon packet emission and diffuse re-emission.
 *
 * Behavioural contract:
 *   PhotonSource::get_random_photon        /root/reference/src/PhotonSource.cpp:208-249
 *   PhotonSource::get_random_direction     /root/reference/src/PhotonSource.hpp:141-148
 *   PhotonSource::set_cross_sections       /root/reference/src/PhotonSource.cpp:189-199
 *   PhotonSource::reemit                   ...:272-308
 *   IsotropicContinuousPhotonSource::get_random_incoming_direction
 *                                          /root/reference/src/IsotropicContinuousPhotonSource.hpp:106-180
 *   MonochromaticPhotonSourceSpectrum::get_random_frequency  MonochromaticPhotonSourceSpectrum.hpp:97-100
 *   PlanckPhotonSourceSpectrum::get_random_frequency         PlanckPhotonSourceSpectrum.cpp:149-165
 *   HydrogenLymanContinuumSpectrum::get_random_frequency     HydrogenLymanContinuumSpectrum.cpp:136-154
 *   HeliumLymanContinuumSpectrum::get_random_frequency       HeliumLymanContinuumSpectrum.cpp:147-165
 *   HeliumTwoPhotonContinuumSpectrum::get_random_frequency   HeliumTwoPhotonContinuumSpectrum.cpp:167-178
 *   PhysicalDiffuseReemissionHandler::reemit(Photon...)      PhysicalDiffuseReemissionHandler.cpp:219-370
 *   PhysicalDiffuseReemissionHandler::set_reemission_probabilities  .hpp:66-106
 *   FixedValueDiffuseReemissionHandler::reemit               FixedValueDiffuseReemissionHandler.hpp:88-101
 *   Utilities::locate                      /root/reference/src/Utilities.hpp:726-742
 *
 * The order in which uniforms are consumed follows the reference exactly (one
 * draw for discrete-vs-continuous even when there is no continuous source, one
 * for the source index, two for the direction, 0/1 for the frequency, one for
 * tau), so that a packet's stream means the same thing in both codes.
 */
#pragma once
#include "cmib_common.cuh"
#include "cross_sections.cuh"
#include "rng.cuh"

namespace cmib {

enum SpectrumKind : int { SPECTRUM_MONOCHROMATIC = 0, SPECTRUM_PLANCK = 1, SPECTRUM_UNIFORM = 2, SPECTRUM_TABULATED = 3 };
enum ReemissionKind : int { REEMISSION_NONE = 0, REEMISSION_PHYSICAL = 1, REEMISSION_FIXED = 2 };
enum ContinuousKind : int { CONTINUOUS_NONE = 0, CONTINUOUS_ISOTROPIC = 1, CONTINUOUS_PLANAR = 2, CONTINUOUS_DISTANT_STAR = 3,
                             CONTINUOUS_EXTENDED_DISC = 4, CONTINUOUS_SPIRAL_GALAXY = 5 };

constexpr int SPECTRUM_NUMFREQ = 1000; /* all tabulated spectra use 1000 frequency bins */
constexpr int LYC_NUMTEMP = 100;

/*
 * A PhotonSourceSpectrum as the device sees it.  MONOCHROMATIC and UNIFORM are closed forms, PLANCK
 * keeps the reference's log-log table; TABULATED is the form every other spectrum of the reference
 * samples from (FaucherGiguere, WMBasic, PopStar, Pegase3, CastelliKurucz, Masked: a frequency grid
 * and its cumulative distribution, inverted with linear interpolation), so any of them runs on the
 * device once its two arrays are handed over (cmib_set_spectrum_table).
 */
struct SpectrumModel {
  int kind;
  double mono_frequency;
  const double *planck;         /* [3][1000]: cdf, log10 cdf, log10 nu/13.6eV */
  const uint16_t *planck_guide; /* bracket guide of the CDF search (locate_guided); NULL = plain bisection */
  int n;                        /* TABULATED: number of frequencies */
  const double *freq, *cdf;     /* TABULATED: [n], [n] */
};

/* everything the emission / re-emission code needs; pointers are device pointers */
constexpr int GALAXY_NBIN = 1000;
/* the constants of SpiralGalaxyContinuousPhotonSource (constructor, .hpp:78-105) */
struct GalaxyModel {
  double rJ, h_stars, bulge_to_total_ratio, rB_over_rJ_plus_rB, rC_over_rJ_plus_rC;
};
/* the constructor's arithmetic: bulge radii 0.2 / 2 / 0.4 kpc, the bulge fraction corrected for the core, and the
 * cumulative disc luminosity 1 - (1 + w / r_stars) exp(-w / r_stars) on 1000 bins up to 1.2 |box anchor| */
inline void build_galaxy_model(const double *box_anchor, double r_stars, double h_stars, double B_over_T, GalaxyModel &m,
                               double *w_table, double *cdf_table) {
  const double kpc = 3.086e19, rC = 0.2 * kpc, rB = 2. * kpc;
  m.rJ = 0.4 * kpc;
  m.h_stars = h_stars;
  m.rB_over_rJ_plus_rB = rB / (rB + m.rJ);
  m.rC_over_rJ_plus_rC = rC / (rC + m.rJ);
  m.bulge_to_total_ratio = B_over_T * (1. - m.rC_over_rJ_plus_rC / m.rB_over_rJ_plus_rB);
  const double rmax = 1.2 * sqrt(box_anchor[0] * box_anchor[0] + box_anchor[1] * box_anchor[1] + box_anchor[2] * box_anchor[2]);
  for (int i = 0; i < GALAXY_NBIN; ++i) {
    const double w = i * rmax / GALAXY_NBIN;
    w_table[i] = w;
    const double x = w / r_stars;
    cdf_table[i] = 1. - (1. + x) * exp(-x);
  }
  w_table[GALAXY_NBIN] = rmax;
  cdf_table[GALAXY_NBIN] = 1.;
}

struct SourceModel {
  /* discrete sources (PhotonSource.cpp:74-100) */
  int n_sources;
  const double *src_pos;   /* [n_sources][3] */
  const double *src_cum;   /* cumulative probabilities, last == 1 */
  /* discrete vs continuous source (PhotonSource.cpp:100-131): probability 0.5 when both exist;
   * packets of the two kinds carry different weights */
  double continuous_probability; /* 0: discrete sources only, 1: continuous source only */
  double discrete_weight;        /* 1 (0 without discrete sources) */
  double continuous_weight;      /* L_continuous / L_discrete (1 without discrete sources) */
  int continuous_kind;
  /* CONTINUOUS_PLANAR (PlanarContinuousPhotonSource.hpp:50-131): the plane coordinate[planar_axis] =
   * planar_intercept, rectangle anchor + [0, sides) in the two other coordinates (ascending index) */
  int planar_axis;
  double planar_intercept, planar_anchor[2], planar_sides[2];
  /* CONTINUOUS_EXTENDED_DISC (ExtendedDiscContinuousPhotonSource.hpp:60-210): emission from the volume of a disc:
   * coordinate[planar_axis] Gaussian around planar_intercept with this scale height, uniform over the box in the
   * two other coordinates */
  double disc_scale_height;
  /* CONTINUOUS_SPIRAL_GALAXY (SpiralGalaxyContinuousPhotonSource.hpp:46-195): a Jaffe-like bulge + a double
   * exponential stellar disc around the origin; galaxy_w / galaxy_cdf = the cumulative radial luminosity of the disc */
  GalaxyModel galaxy;
  const double *galaxy_w, *galaxy_cdf; /* [GALAXY_NBIN + 1] each */
  /* CONTINUOUS_DISTANT_STAR (DistantStarContinuousPhotonSource.hpp:60-90): a star outside the box;
   * star_exposed[d] = -1 / +1 / 0: the star lies below / above / within the box along d */
  double star_position[3];
  int star_exposed[3];
  SpectrumModel cont_spectrum;   /* spectrum of the continuous source */
  SpectrumModel spectrum;        /* spectrum of the discrete sources */
  /* cross sections */
  int xs_kind;
  double xs_fixed[NUM_IONS];     /* FixedValue; Bimodal: the values below the frequency limit */
  double xs_high[NUM_IONS];      /* Bimodal (BimodalCrossSections.hpp:247-254): at and above xs_limit */
  double xs_limit;
  double A_He;
  /* what the packet carries as sigma[ion]: the cross section times fold[ion].  IonizationSimulation: 1 for every
   * ion.  TaskBasedIonizationSimulation folds the abundance of the ion's element into the cross section of every ion
   * but H0 (SourceDiscretePhotonTaskContext.hpp:172-180, PhotonReemitTaskContext.hpp:150-156) and divides the mean
   * intensities by it before the state update (TaskBasedIonizationSimulation.cpp:932-951); its re-emission decision
   * then uses a helium abundance of 1 (PhotonReemitTaskContext.hpp:121-127): A_He_reemit */
  double fold[NUM_IONS];
  double A_He_reemit;
  /* diffuse re-emission */
  int reemission_kind;
  double fixed_reemission_probability;
  double fixed_reemission_frequency;
  const double *hlyc_freq, *hlyc_temp, *hlyc_cdf;    /* [1000], [100], [100][1000] */
  const double *helyc_freq, *helyc_temp, *helyc_cdf; /* same shapes */
  const double *he2pc_freq, *he2pc_cdf;              /* [1000], [1000] */
  /* bracket guides of the CDF searches (locate_guided); NULL = plain bisection */
  const uint16_t *hlyc_guide, *helyc_guide, *he2pc_guide; /* [rows][GUIDE_N + 1] */
};

/* Utilities::locate: bisection, result clamped to [0, length-2] */
CMIB_HD uint32_t locate(double x, const double *xarr, uint32_t length) {
  uint32_t jl = 0, ju = length;
  while (ju - jl > 1) {
    const uint32_t jm = (ju + jl) >> 1;
    if (x > xarr[jm]) jl = jm; else ju = jm;
  }
  if (jl == length - 1) --jl;
  return jl;
}

/*
 * Utilities::locate with a narrowed start bracket.  For a non-decreasing array the result of the
 * bisection does not depend on the probes: it is the largest j with xarr[j] < x, clamped to
 * [0, length-2].  guide[g] = largest j with xarr[j] < g/GUIDE_N (0 if none), tabulated on the
 * host for g = 0..GUIDE_N, brackets every x in [g/GUIDE_N, (g+1)/GUIDE_N): the 10 dependent loads
 * of a 1000-entry search become ~2-3 (the searches were half of the re-emission decision kernel).
 * x must lie in [0, 1] (a cumulative distribution is searched with a uniform deviate).
 */
constexpr int GUIDE_N = 256;
CMIB_HD uint32_t locate_guided(double x, const double *xarr, uint32_t length, const uint16_t *guide) {
  if (guide == nullptr) return locate(x, xarr, length);
  int g = (int)(x * (double)GUIDE_N);
  g = (g < 0) ? 0 : ((g > GUIDE_N - 1) ? GUIDE_N - 1 : g);
  uint32_t jl = guide[g], ju = (uint32_t)guide[g + 1] + 1u;
  if (ju > length) ju = length;
  while (ju - jl > 1) {
    const uint32_t jm = (ju + jl) >> 1;
    if (x > xarr[jm]) jl = jm; else ju = jm;
  }
  if (jl == length - 1) --jl;
  return jl;
}

/* The samplers below are templates over the generator: the device passes the packet's Philox stream
 * (PacketRng, rng.cuh), the CPU-tier tests pass the reference's RANLUX stream (host/RandomGenerator.hpp)
 * and must then reproduce the reference bit for bit.  A generator is anything with rng_uniform(g). */
template <class Rng>
CMIB_HD void random_direction(Rng &rng, double &dx, double &dy, double &dz) {
  const double cost = 2. * rng_uniform(rng) - 1.;
  const double s2 = 1. - cost * cost;
  const double sint = sqrt(s2 > 0. ? s2 : 0.);
  const double phi = 2. * M_PI * rng_uniform(rng);
  double sinp, cosp;
#if defined(__CUDA_ARCH__)
  sincos(phi, &sinp, &cosp);
#else
  sinp = sin(phi);
  cosp = cos(phi);
#endif
  dx = sint * cosp;
  dy = sint * sinp;
  dz = cost;
}

/*
 * IsotropicContinuousPhotonSource::get_random_incoming_direction (.hpp:106-180): a focus point
 * uniform in the box (u[0..2]), an isotropic direction through it (u[3], u[4]); the packet starts
 * where that line enters the box, moved inside the half-open box by at most one epsilon.
 */
CMIB_HD void isotropic_incoming(const GridGeom &g, const double *u, double &px, double &py, double &pz,
                                double &dx, double &dy, double &dz) {
  const double fx = g.anchor[0] + g.sides[0] * u[0];
  const double fy = g.anchor[1] + g.sides[1] * u[1];
  const double fz = g.anchor[2] + g.sides[2] * u[2];
  const double cost = 2. * u[3] - 1.;
  const double s2 = 1. - cost * cost;
  const double sint = sqrt(s2 > 0. ? s2 : 0.);
  const double phi = 2. * M_PI * u[4];
  double sinp, cosp;
#if defined(__CUDA_ARCH__)
  sincos(phi, &sinp, &cosp);
#else
  cosp = cos(phi);
  sinp = sin(phi);
#endif
  dx = sint * cosp;
  dy = sint * sinp;
  dz = cost;
  const double tx = g.anchor[0] + g.sides[0], ty = g.anchor[1] + g.sides[1], tz = g.anchor[2] + g.sides[2];
  const double lx = (dx < 0.) ? (tx - fx) / dx : ((dx > 0.) ? (g.anchor[0] - fx) / dx : -DBL_MAX);
  const double ly = (dy < 0.) ? (ty - fy) / dy : ((dy > 0.) ? (g.anchor[1] - fy) / dy : -DBL_MAX);
  const double lz = (dz < 0.) ? (tz - fz) / dz : ((dz > 0.) ? (g.anchor[2] - fz) / dz : -DBL_MAX);
  const double lxy = (lx < ly) ? ly : lx;     /* std::max(lx, ly) */
  const double maxl = (lxy < lz) ? lz : lxy;  /* std::max(.., lz) */
  px = fx + maxl * dx;
  py = fy + maxl * dy;
  pz = fz + maxl * dz;
  const double eps = DBL_EPSILON;
  const double hx = tx - eps * g.sides[0], hy = ty - eps * g.sides[1], hz = tz - eps * g.sides[2];
  px = (hx < px) ? hx : px; /* std::min(position, top - eps * side) */
  py = (hy < py) ? hy : py;
  pz = (hz < pz) ? hz : pz;
  px = (px < g.anchor[0]) ? g.anchor[0] : px; /* std::max(position, anchor) */
  py = (py < g.anchor[1]) ? g.anchor[1] : py;
  pz = (pz < g.anchor[2]) ? g.anchor[2] : pz;
}

/* PlanarContinuousPhotonSource::get_random_incoming_direction (.hpp:179-204): a point uniform on the
 * rectangle (u[0], u[1]), an isotropic direction (u[2], u[3]) — both half-spaces, as in the reference */
CMIB_HD void planar_incoming(int axis, double intercept, const double *anchor, const double *sides, const double *u,
                             double &px, double &py, double &pz, double &dx, double &dy, double &dz) {
  const int i0 = (axis + 1) % 3, i1 = (axis + 2) % 3;
  const int lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
  double p[3];
  p[lo] = anchor[0] + u[0] * sides[0];
  p[hi] = anchor[1] + u[1] * sides[1];
  p[axis] = intercept;
  px = p[0]; py = p[1]; pz = p[2];
  const double cost = 2. * u[2] - 1.;
  const double s2 = 1. - cost * cost;
  const double sint = sqrt(s2 > 0. ? s2 : 0.);
  const double phi = 2. * M_PI * u[3];
  double sinp, cosp;
#if defined(__CUDA_ARCH__)
  sincos(phi, &sinp, &cosp);
#else
  cosp = cos(phi);
  sinp = sin(phi);
#endif
  dx = sint * cosp;
  dy = sint * sinp;
  dz = cost;
}

/*
 * ExtendedDiscContinuousPhotonSource::get_random_incoming_direction (.hpp:148-197): a point uniform over the
 * box in the two in-plane coordinates (ascending index), a Gaussian height (Box-Muller, one deviate pair per
 * trial, redrawn while it falls outside the box), then an isotropic direction.  `uniform()` supplies the deviates.
 */
template <class Uniform>
CMIB_HD void extended_disc_incoming(const GridGeom &g, int axis, double origin, double scale_height, Uniform &&uniform,
                                    double &px, double &py, double &pz, double &dx, double &dy, double &dz) {
  const int i0 = (axis + 1) % 3, i1 = (axis + 2) % 3;
  const int lo = i0 < i1 ? i0 : i1, hi = i0 < i1 ? i1 : i0;
  double p[3];
  p[lo] = g.anchor[lo] + uniform() * g.sides[lo];
  p[hi] = g.anchor[hi] + uniform() * g.sides[hi];
  const double bottom = g.anchor[axis], top = g.anchor[axis] + g.sides[axis];
  for (int trial = 0; trial < (1 << 24); ++trial) {
    const double rho = scale_height * sqrt(-2. * log(uniform()));
    p[axis] = rho * cos(2. * M_PI * uniform()) + origin;
    if (!(p[axis] < bottom || p[axis] > top)) break;
  }
  px = p[0]; py = p[1]; pz = p[2];
  const double cost = 2. * uniform() - 1.;
  const double s2 = 1. - cost * cost;
  const double sint = sqrt(s2 > 0. ? s2 : 0.);
  const double phi = 2. * M_PI * uniform();
  double sinp, cosp;
#if defined(__CUDA_ARCH__)
  sincos(phi, &sinp, &cosp);
#else
  cosp = cos(phi);
  sinp = sin(phi);
#endif
  dx = sint * cosp;
  dy = sint * sinp;
  dz = cost;
}

/*
 * SpiralGalaxyContinuousPhotonSource::get_random_incoming_direction (.hpp:131-190): positions are drawn from the
 * bulge (with probability bulge_to_total_ratio: r = rJ / (1 / A - 1), A uniform between the two bulge constants,
 * isotropic) or from the disc (exponential height, radius from the tabulated cumulative luminosity) until one
 * lies inside the box (Box::inside: anchor <= x < anchor + sides), then an isotropic direction.
 */
template <class Uniform>
CMIB_HD void spiral_galaxy_incoming(const GridGeom &g, const GalaxyModel &m, const double *w_table, const double *cdf_table,
                                    Uniform &&uniform, double &px, double &py, double &pz, double &dx, double &dy, double &dz) {
  double p[3] = {g.anchor[0] - g.sides[0], g.anchor[1] - g.sides[1], g.anchor[2] - g.sides[2]};
  for (int trial = 0; trial < (1 << 24); ++trial) {
    const double x_bulge = uniform();
    if (x_bulge <= m.bulge_to_total_ratio) {
      const double u = uniform();
      const double A = u * m.rB_over_rJ_plus_rB + (1. - u) * m.rC_over_rJ_plus_rC;
      const double r = m.rJ / (1. / A - 1.);
      const double phi = 2. * M_PI * uniform();
      const double cost = 2. * uniform() - 1.;
      const double s2 = 1. - cost * cost;
      const double sint = sqrt(s2 > 0. ? s2 : 0.);
      double sinp, cosp;
#if defined(__CUDA_ARCH__)
      sincos(phi, &sinp, &cosp);
#else
      cosp = cos(phi);
      sinp = sin(phi);
#endif
      p[0] = r * sint * cosp;
      p[1] = r * sint * sinp;
      p[2] = r * cost;
    } else {
      const double u1 = 2. * uniform() - 1.;
      const double z = (u1 > 0.) ? -m.h_stars * log(u1) : m.h_stars * log(-u1);
      const double phi = 2. * M_PI * uniform();
      const double u2 = uniform();
      const uint32_t i = locate(u2, cdf_table, GALAXY_NBIN + 1);
      const double w = w_table[i] + (u2 - cdf_table[i]) / (cdf_table[i + 1] - cdf_table[i]) * (w_table[i + 1] - w_table[i]);
      double sinp, cosp;
#if defined(__CUDA_ARCH__)
      sincos(phi, &sinp, &cosp);
#else
      cosp = cos(phi);
      sinp = sin(phi);
#endif
      p[0] = w * cosp;
      p[1] = w * sinp;
      p[2] = z;
    }
    if (p[0] >= g.anchor[0] && p[0] < g.anchor[0] + g.sides[0] && p[1] >= g.anchor[1] && p[1] < g.anchor[1] + g.sides[1] &&
        p[2] >= g.anchor[2] && p[2] < g.anchor[2] + g.sides[2])
      break;
  }
  px = p[0]; py = p[1]; pz = p[2];
  const double cost = 2. * uniform() - 1.;
  const double s2 = 1. - cost * cost;
  const double sint = sqrt(s2 > 0. ? s2 : 0.);
  const double phi = 2. * M_PI * uniform();
  double sinp, cosp;
#if defined(__CUDA_ARCH__)
  sincos(phi, &sinp, &cosp);
#else
  cosp = cos(phi);
  sinp = sin(phi);
#endif
  dx = sint * cosp;
  dy = sint * sinp;
  dz = cost;
}

/*
 * DistantStarContinuousPhotonSource::get_random_incoming_direction (.hpp:164-192) with enters_box
 * (:125-155): isotropic directions from the star (the first one mirrored towards the box) are drawn
 * until one hits an exposed face within the face's bounds; the packet starts at that intersection
 * point.  `uniform()` supplies the deviates (two per trial).  Kept as in the reference: only the
 * first trial is mirrored; the start is exactly ON the face, so for a face at the top anchor
 * get_cell_indices can put it one cell outside and the packet is lost (Isotropic guards against
 * that with its epsilon, this source does not).
 */
template <class Uniform>
CMIB_HD void distant_star_incoming(const GridGeom &g, const double *star, const int *exposed, Uniform &&uniform,
                                   double &px, double &py, double &pz, double &dx, double &dy, double &dz) {
  const double bottom[3] = {g.anchor[0], g.anchor[1], g.anchor[2]};
  const double top[3] = {g.anchor[0] + g.sides[0], g.anchor[1] + g.sides[1], g.anchor[2] + g.sides[2]};
  double d[3], fp[3] = {0., 0., 0.};
  for (int trial = 0; trial < (1 << 24); ++trial) {
    const double cost = 2. * uniform() - 1.;
    const double s2 = 1. - cost * cost;
    const double sint = sqrt(s2 > 0. ? s2 : 0.);
    const double phi = 2. * M_PI * uniform();
    double sinp, cosp;
#if defined(__CUDA_ARCH__)
    sincos(phi, &sinp, &cosp);
#else
    cosp = cos(phi);
    sinp = sin(phi);
#endif
    d[0] = sint * cosp; d[1] = sint * sinp; d[2] = cost;
    if (trial == 0)
      for (int i = 0; i < 3; ++i)
        if (exposed[i] * d[i] > 0.) d[i] = -d[i];
    bool enters = false;
    for (int i = 0; i < 3 && !enters; ++i) {
      if (exposed[i] * d[i] < 0.) {
        const double plane = (exposed[i] < 0) ? bottom[i] : top[i];
        const double l = (plane - star[i]) / d[i];
        for (int k = 0; k < 3; ++k) fp[k] = star[k] + l * d[k];
        const int j1 = (i + 1) % 3, j2 = (i + 2) % 3;
        enters = fp[j1] >= bottom[j1] && fp[j1] <= top[j1] && fp[j2] >= bottom[j2] && fp[j2] <= top[j2];
      }
    }
    if (enters) break;
  }
  px = fp[0]; py = fp[1]; pz = fp[2];
  dx = d[0]; dy = d[1]; dz = d[2];
}

/* PlanckPhotonSourceSpectrum::get_random_frequency (PlanckPhotonSourceSpectrum.cpp:149-165) for the deviate x */
CMIB_HD double planck_frequency_at(const double *tab, double x, const uint16_t *guide = nullptr) {
  const double *cdf = tab, *logcdf = tab + SPECTRUM_NUMFREQ, *lognu = tab + 2 * SPECTRUM_NUMFREQ;
  const uint32_t ix = locate_guided(x, cdf, SPECTRUM_NUMFREQ, guide);
  const double lf = (log10(x) - logcdf[ix]) / (logcdf[ix + 1] - logcdf[ix]) *
                        (lognu[ix + 1] - lognu[ix]) + lognu[ix];
  return fpow(10., lf) * 3.288465385e15; /* device: exp(lf ln 10), cmib_common.cuh */
}
template <class Rng>
CMIB_HD double planck_frequency(const double *tab, Rng &rng, const uint16_t *guide = nullptr) {
  return planck_frequency_at(tab, rng_uniform(rng), guide);
}

/* PhotonSourceSpectrum::get_random_frequency of the source spectra:
 *   Monochromatic  MonochromaticPhotonSourceSpectrum.hpp:97-100 (no deviate consumed)
 *   Planck         PlanckPhotonSourceSpectrum.cpp:149-165
 *   Uniform        UniformPhotonSourceSpectrum.hpp:50-53
 *   tabulated      e.g. FaucherGiguerePhotonSourceSpectrum.cpp:234-247, WMBasicPhotonSourceSpectrum.cpp:236-248,
 *                  MaskedPhotonSourceSpectrum.cpp:123-135 */
CMIB_HD double uniform_frequency(double x) { return (1. + 3. * x) * 3.289e15; }
CMIB_HD double tabulated_frequency(const double *freq, const double *cdf, uint32_t n, double x) {
  const uint32_t inu = locate(x, cdf, n);
  return freq[inu] + (freq[inu + 1] - freq[inu]) * (x - cdf[inu]) / (cdf[inu + 1] - cdf[inu]);
}
template <class Rng>
CMIB_HD double spectrum_frequency(const SpectrumModel &sp, Rng &rng) {
  if (sp.kind == SPECTRUM_PLANCK) return planck_frequency(sp.planck, rng, sp.planck_guide);
  if (sp.kind == SPECTRUM_UNIFORM) return uniform_frequency(rng_uniform(rng));
  if (sp.kind == SPECTRUM_TABULATED) return tabulated_frequency(sp.freq, sp.cdf, (uint32_t)sp.n, rng_uniform(rng));
  return sp.mono_frequency;
}

/* Utilities::locate on two arrays of the same length at once: the two bisections are independent,
 * running them in lock step puts their (L2-latency bound) loads in flight together */
CMIB_HD void locate2(double x, const double *a, const double *b, uint32_t length, uint32_t &ja, uint32_t &jb) {
  uint32_t la = 0, ua = length, lb = 0, ub = length;
  while (ua - la > 1 || ub - lb > 1) {
    const uint32_t ma = (ua + la) >> 1, mb = (ub + lb) >> 1;
    const double va = a[ma], vb = b[mb];
    if (ua - la > 1) { if (x > va) la = ma; else ua = ma; }
    if (ub - lb > 1) { if (x > vb) lb = mb; else ub = mb; }
  }
  if (la == length - 1) --la;
  if (lb == length - 1) --lb;
  ja = la;
  jb = lb;
}

template <class Rng>
CMIB_HD double lyc_frequency(const double *freq, const double *temp, const double *cdf, double T,
                             Rng &rng, const uint16_t *guide = nullptr) {
  const uint32_t iT = locate(T, temp, LYC_NUMTEMP);
  const double x = rng_uniform(rng);
  uint32_t inu1, inu2;
  if (guide) {
    inu1 = locate_guided(x, cdf + (size_t)iT * SPECTRUM_NUMFREQ, SPECTRUM_NUMFREQ, guide + (size_t)iT * (GUIDE_N + 1));
    inu2 = locate_guided(x, cdf + (size_t)(iT + 1) * SPECTRUM_NUMFREQ, SPECTRUM_NUMFREQ,
                         guide + (size_t)(iT + 1) * (GUIDE_N + 1));
  } else {
    locate2(x, cdf + (size_t)iT * SPECTRUM_NUMFREQ, cdf + (size_t)(iT + 1) * SPECTRUM_NUMFREQ, SPECTRUM_NUMFREQ,
            inu1, inu2);
  }
  return freq[inu1] + (T - temp[iT]) * (freq[inu2] - freq[inu1]) / (temp[iT + 1] - temp[iT]);
}

template <class Rng>
CMIB_HD double he2pc_frequency(const double *freq, const double *cdf, Rng &rng,
                               const uint16_t *guide = nullptr) {
  const double x = rng_uniform(rng);
  const uint32_t inu = locate_guided(x, cdf, SPECTRUM_NUMFREQ, guide);
  return freq[inu] + (freq[inu + 1] - freq[inu]) * (x - cdf[inu]) / (cdf[inu + 1] - cdf[inu]);
}

/*
 * PhotonSource::get_random_photon (PhotonSource.cpp:208-249) up to the cross sections: one draw
 * decides between the discrete sources (source index, isotropic direction, frequency from their
 * spectrum) and the continuous source (position + direction on the box surface, frequency from its
 * own spectrum).  isrc = index of the discrete source, -1 for a packet of the continuous source.
 */
/* the continuous-source branch of get_random_photon, kept out of line on the device: it is rarely taken
 * and its code (a rejection loop, three geometries) must not set the register budget of prepare_kernel */
#if defined(__CUDACC__)
#define CMIB_HD_OUT_OF_LINE __host__ __device__ __noinline__
#else
#define CMIB_HD_OUT_OF_LINE inline
#endif
template <class Rng>
CMIB_HD_OUT_OF_LINE void emit_continuous(const SourceModel &m, const GridGeom &g, Rng &rng, double &px, double &py, double &pz,
                                         double &dx, double &dy, double &dz, double &nu) {
    double u[5];
    if (m.continuous_kind == CONTINUOUS_DISTANT_STAR) {
      distant_star_incoming(g, m.star_position, m.star_exposed, [&rng]() { return rng_uniform(rng); }, px, py, pz, dx, dy, dz);
    } else if (m.continuous_kind == CONTINUOUS_SPIRAL_GALAXY) {
      spiral_galaxy_incoming(g, m.galaxy, m.galaxy_w, m.galaxy_cdf, [&rng]() { return rng_uniform(rng); }, px, py, pz, dx, dy, dz);
    } else if (m.continuous_kind == CONTINUOUS_EXTENDED_DISC) {
      extended_disc_incoming(g, m.planar_axis, m.planar_intercept, m.disc_scale_height, [&rng]() { return rng_uniform(rng); },
                             px, py, pz, dx, dy, dz);
    } else if (m.continuous_kind == CONTINUOUS_PLANAR) {
#pragma unroll
      for (int k = 0; k < 4; ++k) u[k] = rng_uniform(rng);
      planar_incoming(m.planar_axis, m.planar_intercept, m.planar_anchor, m.planar_sides, u, px, py, pz, dx, dy, dz);
    } else {
#pragma unroll
      for (int k = 0; k < 5; ++k) u[k] = rng_uniform(rng);
      isotropic_incoming(g, u, px, py, pz, dx, dy, dz);
    }
  nu = spectrum_frequency(m.cont_spectrum, rng);
}

template <class Rng>
CMIB_HD void emit_primary(const SourceModel &m, const GridGeom &g, Rng &rng, double &px, double &py,
                          double &pz, double &dx, double &dy, double &dz, double &nu, int &isrc) {
  double x = rng_uniform(rng);
  if (x >= m.continuous_probability) {
    x = rng_uniform(rng);
    isrc = 0;
    while (isrc < m.n_sources - 1 && x > m.src_cum[isrc]) ++isrc;
    px = m.src_pos[3 * isrc];
    py = m.src_pos[3 * isrc + 1];
    pz = m.src_pos[3 * isrc + 2];
    random_direction(rng, dx, dy, dz);
    nu = spectrum_frequency(m.spectrum, rng);
  } else {
    emit_continuous(m, g, rng, px, py, pz, dx, dy, dz, nu);
    isrc = -1;
  }
}

/* the five cumulative re-emission probabilities of a cell at temperature T */
CMIB_HD void reemission_probabilities(double T, double *p) {
  const double T4 = T * 1.e-4;
  const double alpha_1_H = 1.58e-13 * fpow(T4, -0.53);
  const double alpha_A_agn = 4.18e-13 * fpow(T4, -0.7);
  p[REEMIT_H] = alpha_1_H / alpha_A_agn;
  const double alpha_1_He = 1.54e-13 * fpow(T4, -0.486);
  const double alpha_e_2tS = 2.1e-13 * fpow(T4, -0.381);
  const double alpha_e_2sS = 2.06e-14 * fpow(T4, -0.451);
  const double alpha_e_2sP = 4.17e-14 * fpow(T4, -0.695);
  const double alphaHe = alpha_1_He + alpha_e_2tS + alpha_e_2sS + alpha_e_2sP;
  const double He_LyC = alpha_1_He / alphaHe;
  const double He_NpEEv = He_LyC + alpha_e_2tS / alphaHe;
  const double He_TPC = He_NpEEv + alpha_e_2sS / alphaHe;
  const double He_LyA = He_TPC + alpha_e_2sP / alphaHe;
  p[REEMIT_HE_LYC] = He_LyC;
  p[REEMIT_HE_NPEEV] = He_NpEEv;
  p[REEMIT_HE_TPC] = He_TPC;
  p[REEMIT_HE_LYA] = He_LyA;
}

/* cross sections of a packet at frequency nu.  NSIG = 14 (all ions) or 1 (H only) */
template <int NSIG>
CMIB_HD void packet_cross_sections(const SourceModel &m, double nu, double *sigma,
                                   double &sigma_He_corr) {
  if (m.xs_kind == XS_VERNER) {
    double sHe = 0.;
#pragma unroll 1
    for (int ion = 0; ion < NUM_IONS; ++ion) {
      const double s = verner_cross_section(ion, nu);
      if (ion < NSIG) sigma[ion] = s * m.fold[ion]; /* x * 1 == x: the legacy convention is untouched */
      if (ion == ION_He_n) sHe = s;
    }
    sigma_He_corr = m.A_He * sHe;
  } else if (m.xs_kind == XS_BIMODAL) {
    const bool low = nu < m.xs_limit;
#pragma unroll
    for (int ion = 0; ion < NSIG; ++ion) sigma[ion] = (low ? m.xs_fixed[ion] : m.xs_high[ion]) * m.fold[ion];
    sigma_He_corr = m.A_He * (low ? m.xs_fixed[ION_He_n] : m.xs_high[ION_He_n]);
  } else {
#pragma unroll
    for (int ion = 0; ion < NSIG; ++ion) sigma[ion] = m.xs_fixed[ion] * m.fold[ion];
    sigma_He_corr = m.A_He * m.xs_fixed[ION_He_n];
  }
}

/*
 * Physical diffuse re-emission decision.  xH, xHe, T = state of the absorbing
 * cell, p = its 5 cumulative probabilities.  Returns the new frequency (0 =
 * packet absorbed) and the new packet type.
 */
template <class Rng>
CMIB_HD double physical_reemit(const SourceModel &m, double sigma_H, double sigma_He, double xH,
                               double xHe, double T, const double *p, Rng &rng, int &type) {
  double nu = 0.;
  const double nH0anuH0 = xH * sigma_H;
  const double nHe0anuHe0 = xHe * m.A_He_reemit * sigma_He;
  const double pHabs = nH0anuH0 / (nH0anuH0 + nHe0anuHe0);
  double x = rng_uniform(rng);
  type = PACKET_ABSORBED;
  if (x <= pHabs) {
    x = rng_uniform(rng);
    if (x <= p[REEMIT_H]) {
      nu = lyc_frequency(m.hlyc_freq, m.hlyc_temp, m.hlyc_cdf, T, rng, m.hlyc_guide);
      type = PACKET_DIFFUSE_HI;
    }
  } else {
    x = rng_uniform(rng);
    if (x <= p[REEMIT_HE_LYC]) {
      nu = lyc_frequency(m.helyc_freq, m.helyc_temp, m.helyc_cdf, T, rng, m.helyc_guide);
      type = PACKET_DIFFUSE_HeI;
    } else if (x <= p[REEMIT_HE_NPEEV]) {
      nu = 4.788e15;
      type = PACKET_DIFFUSE_HeI;
    } else if (x <= p[REEMIT_HE_TPC]) {
      x = rng_uniform(rng);
      if (x < 0.56) {
        nu = he2pc_frequency(m.he2pc_freq, m.he2pc_cdf, rng, m.he2pc_guide);
        type = PACKET_DIFFUSE_HeI;
      }
    } else if (x <= p[REEMIT_HE_LYA]) {
      const double sqrtTnH0 = sqrt(T) * xH;
      const double pHots = sqrtTnH0 / (sqrtTnH0 + 77. * xHe);
      x = rng_uniform(rng);
      if (x < pHots) {
        x = rng_uniform(rng);
        if (x <= p[REEMIT_H]) {
          nu = lyc_frequency(m.hlyc_freq, m.hlyc_temp, m.hlyc_cdf, T, rng, m.hlyc_guide);
          type = PACKET_DIFFUSE_HI;
        }
      } else {
        x = rng_uniform(rng);
        if (x < 0.56) {
          nu = he2pc_frequency(m.he2pc_freq, m.he2pc_cdf, rng, m.he2pc_guide);
          type = PACKET_DIFFUSE_HeI;
        }
      }
    }
  }
  return nu;
}

} // namespace cmib
