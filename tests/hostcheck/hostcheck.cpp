/*
 * tests/hostcheck/hostcheck.cpp — TEST INFRASTRUCTURE ONLY.
 *
 * Compiles the product's __host__ __device__ physics headers
 * (cmacionize_b200/csrc/*.cuh) with plain g++ so that their logic can be
 * checked against the oracle and the reference's golden vectors on machines
 * without a GPU (the `-m "not gpu"` tier).  It is NOT a CPU fallback: nothing
 * under cmacionize_b200/ loads this library, and the product's entry points
 * fail when no B200 is present.  The GPU tier (`-m gpu`) repeats every one of
 * these checks through the real C ABI on the device.
 */
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../cmacionize_b200/host/RandomGenerator.hpp"
#include "../../cmacionize_b200/csrc/march.cuh"
#include "../../cmacionize_b200/csrc/shoot.cuh"
#include "../../cmacionize_b200/csrc/source.cuh"
#include "../../cmacionize_b200/csrc/spectrum_tables.hpp"
#include "../../cmacionize_b200/csrc/state.cuh"

/* the reference's RANLUX stream as a generator for the samplers of source.cuh (rng_uniform is found
 * by argument-dependent lookup when the templates are instantiated) */
struct RanluxRng {
  cmi::RandomGenerator g;
  explicit RanluxRng(int seed) : g(seed) {}
};
inline double rng_uniform(RanluxRng &r) { return r.g.get_uniform_random_double(); }

using namespace cmib;

extern "C" {

void hc_verner_cross_sections(int64_t n, const double *nu, double *sigma) {
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < NUM_IONS; ++k) sigma[i * NUM_IONS + k] = verner_cross_section(k, nu[i]);
}

void hc_verner_recombination_rates(int64_t n, const double *T, double *alpha) {
  for (int64_t i = 0; i < n; ++i)
    for (int k = 0; k < NUM_IONS; ++k) alpha[i * NUM_IONS + k] = verner_recombination_rate(k, T[i]);
}

void hc_charge_transfer(int64_t n, const double *T4, double *out) {
  for (int64_t i = 0; i < n; ++i) {
    double *o = out + i * 3 * NUM_IONS;
    for (int k = 0; k < NUM_IONS; ++k) {
      o[k] = (k == ION_H_n) ? 0. : ct_recombination_H(k, T4[i]);
      o[NUM_IONS + k] = ct_ionization_H(k, T4[i]);
      o[2 * NUM_IONS + k] = ct_recombination_He(k, T4[i]);
    }
  }
}

void hc_line_cooling(int64_t n, const double *T, const double *ne, const double *abund, double *c) {
  for (int64_t i = 0; i < n; ++i) c[i] = line_cooling(T[i], ne[i], abund + i * LC_NUM);
}

void hc_solve5(int64_t n, double *A, double *B, int32_t *status) {
  for (int64_t i = 0; i < n; ++i) {
    double a[5][5];
    memcpy(a, A + 25 * i, sizeof(a));
    status[i] = solve5(a, B + 5 * i);
    memcpy(A + 25 * i, a, sizeof(a));
  }
}

void hc_reemission_probabilities(int64_t n, const double *T, double *out) {
  for (int64_t i = 0; i < n; ++i) reemission_probabilities(T[i], out + i * NUM_REEMIT);
}

static RecombinationModel make_rr(int kind, const double *fixed) {
  RecombinationModel rr;
  rr.kind = kind;
  for (int k = 0; k < NUM_IONS; ++k) rr.fixed[k] = fixed ? fixed[k] : 0.;
  return rr;
}

void hc_ionization_state(int64_t n, double jfac, double hfac, const double *abund, int rr_kind,
                         const double *rr_fixed, const double *J, const double *heat,
                         const double *ndens, const double *T, double *x, double *heat_out) {
  const RecombinationModel rr = make_rr(rr_kind, rr_fixed);
  for (int64_t i = 0; i < n; ++i) {
    double j[NUM_IONS], h[2];
    for (int k = 0; k < NUM_IONS; ++k) j[k] = J[k * n + i];
    h[0] = heat[i];
    h[1] = heat[n + i];
    CellState out;
    cell_ionization_state(jfac, hfac, j, h, ndens[i], T[i], abund, rr, out);
    for (int k = 0; k < NUM_IONS; ++k) x[k * n + i] = out.x[k];
    heat_out[i] = out.heat[0];
    heat_out[n + i] = out.heat[1];
  }
}

void hc_h_he_state(int64_t n, const double *alphaH, const double *alphaHe, const double *jH,
                   const double *jHe, const double *nH, const double *AHe, const double *T,
                   double *h0, double *he0) {
  for (int64_t i = 0; i < n; ++i)
    ionization_states_hydrogen_helium(alphaH[i], alphaHe[i], jH[i], jHe[i], nH[i], AHe[i], T[i],
                                      h0[i], he0[i]);
}

void hc_cooling_heating_balance(int64_t n, const double *T, const double *ndens, const double *j,
                                const double *h, const double *abund, double pahfac, double crfac,
                                double crscale, const double *midz, int rr_kind,
                                const double *rr_fixed, double *h0, double *he0, double *gain,
                                double *loss, double *metals) {
  const RecombinationModel rr = make_rr(rr_kind, rr_fixed);
  for (int64_t i = 0; i < n; ++i) {
    double x[NUM_IONS] = {0.};
    cooling_heating_balance(h0[i], he0[i], gain[i], loss[i], T[i], ndens[i], midz ? midz[i] : 0.,
                            j + i * NUM_IONS, abund, h + 2 * i, pahfac, crfac, crscale, rr, x);
    for (int k = 0; k < 12; ++k) metals[i * 12 + k] = x[2 + k];
  }
}

void hc_temperature(int64_t n, double jfac, double hfac, const double *abund, int rr_kind,
                    const double *rr_fixed, const double *tparams, const double *J,
                    const double *heat, const double *ndens, const double *T,
                    const double *cr_factor, const double *midz, double *T_out, double *x,
                    double *heat_out) {
  const RecombinationModel rr = make_rr(rr_kind, rr_fixed);
  TemperatureParams tp;
  tp.do_temperature = 1;
  tp.min_iterations = 0;
  tp.pahfac = tparams[0];
  tp.crfac = tparams[1];
  tp.crlim = tparams[2];
  tp.crscale = tparams[3];
  tp.min_ionized_T = tparams[4];
  tp.epsilon = tparams[5];
  tp.max_iterations = (uint32_t)tparams[6];
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i) {
    double j[NUM_IONS], h[2], xprev[NUM_IONS] = {0.};
    for (int k = 0; k < NUM_IONS; ++k) j[k] = J[k * n + i];
    h[0] = heat[i];
    h[1] = heat[n + i];
    CellState out;
    cell_temperature(jfac, hfac, j, h, ndens[i], T[i], cr_factor ? cr_factor[i] : -1.,
                     midz ? midz[i] : 0., abund, rr, tp, xprev, out);
    T_out[i] = out.T;
    for (int k = 0; k < NUM_IONS; ++k) x[k * n + i] = out.x[k];
    heat_out[i] = out.heat[0];
    heat_out[n + i] = out.heat[1];
  }
}

void hc_isotropic_incoming(const double *anchor, const double *sides, int64_t n, const double *uniforms,
                           double *pos, double *dir) {
  GridGeom g;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) { g.anchor[d] = anchor[d]; g.sides[d] = sides[d]; }
  for (int64_t i = 0; i < n; ++i)
    isotropic_incoming(g, uniforms + 5 * i, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i], dir[3 * i + 1],
                       dir[3 * i + 2]);
}

void hc_integrate_optical_depth(const double *anchor, const double *sides, const int32_t *ncell, const int32_t *periodic,
                                const double *cell_n, const double *cell_xH, const double *cell_xHe, int64_t np,
                                const double *pos, const double *dir, const double *sigma_H, const double *sigma_He_corr,
                                double *out) {
  GridGeom g;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) {
    g.anchor[d] = anchor[d]; g.sides[d] = sides[d]; g.ncell[d] = ncell[d]; g.periodic[d] = periodic[d];
    g.cellside[d] = sides[d] / ncell[d];
    g.inv_cellside[d] = 1. / g.cellside[d];
  }
  g.ncells = (int64_t)ncell[0] * ncell[1] * ncell[2];
  for (int64_t p = 0; p < np; ++p) {
    MarchState s;
    s.px = pos[3 * p]; s.py = pos[3 * p + 1]; s.pz = pos[3 * p + 2];
    s.dx = dir[3 * p]; s.dy = dir[3 * p + 1]; s.dz = dir[3 * p + 2];
    out[p] = integrate_optical_depth(g, s, sigma_H[p], sigma_He_corr[p], [&](int64_t c) {
      CellOpacity r;
      r.n = cell_n[c]; r.xH = cell_xH[c]; r.xHe = cell_xHe[c]; r.T = 0.;
      return r;
    }, 1ll << 22);
  }
}

/* deviates from the host layer's RANLUX generator (the reference's stream, bit for bit) */
void hc_distant_star_incoming(const double *anchor, const double *sides, const double *star, int seed, int64_t n,
                              double *pos, double *dir) {
  GridGeom g;
  memset(&g, 0, sizeof(g));
  int exposed[3];
  for (int d = 0; d < 3; ++d) {
    g.anchor[d] = anchor[d]; g.sides[d] = sides[d];
    exposed[d] = (star[d] < anchor[d]) ? -1 : ((star[d] > anchor[d] + sides[d]) ? 1 : 0);
  }
  cmi::RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i)
    distant_star_incoming(g, star, exposed, [&rg]() { return rg.get_uniform_random_double(); }, pos[3 * i], pos[3 * i + 1],
                          pos[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
}

void hc_extended_disc_incoming(const double *anchor, const double *sides, int axis, double origin, double scale_height,
                               int seed, int64_t n, double *pos, double *dir) {
  GridGeom g;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) { g.anchor[d] = anchor[d]; g.sides[d] = sides[d]; }
  cmi::RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i)
    extended_disc_incoming(g, axis, origin, scale_height, [&rg]() { return rg.get_uniform_random_double(); }, pos[3 * i],
                           pos[3 * i + 1], pos[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
}

void hc_spiral_galaxy_incoming(const double *anchor, const double *sides, double r_stars, double h_stars, double B_over_T,
                               int seed, int64_t n, double *pos, double *dir, double *tables) {
  GridGeom g;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) { g.anchor[d] = anchor[d]; g.sides[d] = sides[d]; }
  GalaxyModel m;
  build_galaxy_model(anchor, r_stars, h_stars, B_over_T, m, tables, tables + GALAXY_NBIN + 1);
  cmi::RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i)
    spiral_galaxy_incoming(g, m, tables, tables + GALAXY_NBIN + 1, [&rg]() { return rg.get_uniform_random_double(); },
                           pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
}

/* PhotonSource::get_random_photon with the RANLUX stream: n_sources discrete sources with a Planck (T) or
 * monochromatic (nu) spectrum, optionally an isotropic continuous source with its own Planck spectrum
 * (continuous_luminosity > 0), Verner cross sections.  Outputs per packet: pos, dir, nu, sigma[14],
 * A_He sigma_He, weight. */
void hc_random_photons(const double *anchor, const double *sides, int n_sources, const double *src_pos,
                       const double *src_weights, double discrete_luminosity, int spectrum_kind, double spectrum_param,
                       double continuous_luminosity, double continuous_temperature, double A_He, int seed, int64_t n,
                       double *pos, double *dir, double *nu, double *sigma, double *she, double *weight) {
  GridGeom g;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) { g.anchor[d] = anchor[d]; g.sides[d] = sides[d]; }
  SourceModel m;
  memset(&m, 0, sizeof(m));
  m.n_sources = n_sources;
  std::vector<double> cum(n_sources > 0 ? n_sources : 1, 1.);
  for (int i = 0; i < n_sources; ++i) cum[i] = (i ? cum[i - 1] : 0.) + src_weights[i];
  if (n_sources > 0) cum[n_sources - 1] = 1.;
  m.src_pos = src_pos;
  m.src_cum = cum.data();
  std::vector<double> planck, cplanck;
  m.spectrum.kind = spectrum_kind;
  if (spectrum_kind == SPECTRUM_PLANCK) {
    host::build_planck_table(spectrum_param, planck);
    m.spectrum.planck = planck.data();
  } else {
    m.spectrum.mono_frequency = spectrum_param;
  }
  /* PhotonSource.cpp:110-131 */
  if (continuous_luminosity > 0.) {
    m.continuous_kind = CONTINUOUS_ISOTROPIC;
    host::build_planck_table(continuous_temperature, cplanck);
    m.cont_spectrum.kind = SPECTRUM_PLANCK;
    m.cont_spectrum.planck = cplanck.data();
    if (discrete_luminosity > 0.) {
      m.continuous_probability = 0.5;
      m.discrete_weight = 1.;
      m.continuous_weight = (1. - m.continuous_probability) * continuous_luminosity / m.continuous_probability / discrete_luminosity;
    } else {
      m.continuous_probability = 1.;
      m.discrete_weight = 0.;
      m.continuous_weight = 1.;
    }
  } else {
    m.discrete_weight = 1.;
  }
  m.xs_kind = XS_VERNER;
  m.A_He = A_He;
  m.A_He_reemit = A_He;
  for (int k = 0; k < NUM_IONS; ++k) m.fold[k] = 1.; /* IonizationSimulation conventions */
  RanluxRng rng(seed);
  for (int64_t i = 0; i < n; ++i) {
    int isrc;
    emit_primary(m, g, rng, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2], dir[3 * i], dir[3 * i + 1], dir[3 * i + 2], nu[i], isrc);
    packet_cross_sections<NUM_IONS>(m, nu[i], sigma + i * NUM_IONS, she[i]);
    weight[i] = (isrc >= 0) ? m.discrete_weight : m.continuous_weight;
  }
}

/* PhotonSource::reemit with the Physical handler and the RANLUX stream, one call per entry: the packet was
 * absorbed at frequency nu_in[i] in a cell with (xH, xHe, T)[i].  Outputs: new frequency (0 = absorbed for
 * good), packet type, new direction (zeros when absorbed). */
void hc_reemit_sequence(double A_He, int seed, int64_t n, const double *xH, const double *xHe, const double *T,
                        const double *nu_in, double *nu_out, int32_t *type_out, double *dir) {
  SourceModel m;
  memset(&m, 0, sizeof(m));
  m.xs_kind = XS_VERNER;
  m.A_He = A_He;
  m.A_He_reemit = A_He;
  for (int k = 0; k < NUM_IONS; ++k) m.fold[k] = 1.; /* IonizationSimulation conventions */
  m.reemission_kind = REEMISSION_PHYSICAL;
  std::vector<double> hf, ht, hc, hef, het, hec, tf, tc;
  host::build_lyc_table(0, [](double nu) { return verner_cross_section(ION_H_n, nu); }, hf, ht, hc);
  host::build_lyc_table(1, [](double nu) { return verner_cross_section(ION_He_n, nu); }, hef, het, hec);
  host::build_he2pc_table(tf, tc);
  m.hlyc_freq = hf.data(); m.hlyc_temp = ht.data(); m.hlyc_cdf = hc.data();
  m.helyc_freq = hef.data(); m.helyc_temp = het.data(); m.helyc_cdf = hec.data();
  m.he2pc_freq = tf.data(); m.he2pc_cdf = tc.data();
  RanluxRng rng(seed);
  for (int64_t i = 0; i < n; ++i) {
    double sigma[NUM_IONS], she, p[NUM_REEMIT];
    packet_cross_sections<NUM_IONS>(m, nu_in[i], sigma, she);
    reemission_probabilities(T[i], p);
    int type = PACKET_ABSORBED;
    nu_out[i] = physical_reemit(m, sigma[ION_H_n], sigma[ION_He_n], xH[i], xHe[i], T[i], p, rng, type);
    type_out[i] = type;
    dir[3 * i] = dir[3 * i + 1] = dir[3 * i + 2] = 0.;
    if (nu_out[i] != 0.) random_direction(rng, dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
  }
}

void hc_planar_incoming(int axis, double intercept, const double *anchor, const double *sides, int64_t n,
                        const double *uniforms, double *pos, double *dir) {
  for (int64_t i = 0; i < n; ++i)
    planar_incoming(axis, intercept, anchor, sides, uniforms + 4 * i, pos[3 * i], pos[3 * i + 1], pos[3 * i + 2],
                    dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
}

/* tabulated (freq != NULL) or Uniform spectrum sampled with the given deviates */
void hc_tabulated_frequency(int32_t m, const double *freq, const double *cdf, int64_t n, const double *x, double *nu) {
  for (int64_t i = 0; i < n; ++i) nu[i] = freq ? tabulated_frequency(freq, cdf, (uint32_t)m, x[i]) : uniform_frequency(x[i]);
}

void hc_planck_tables(double temperature, double *out) {
  std::vector<double> t;
  host::build_planck_table(temperature, t);
  memcpy(out, t.data(), t.size() * sizeof(double));
}

void hc_lyc_tables(int which, int xs_kind, const double *xs_fixed, double *freq, double *temp,
                   double *cdf) {
  std::vector<double> f, t, c;
  const int ion = which == 0 ? ION_H_n : ION_He_n;
  host::build_lyc_table(which,
                        [=](double nu) {
                          return xs_kind == XS_VERNER ? verner_cross_section(ion, nu) : xs_fixed[ion];
                        },
                        f, t, c);
  memcpy(freq, f.data(), f.size() * sizeof(double));
  memcpy(temp, t.data(), t.size() * sizeof(double));
  memcpy(cdf, c.data(), c.size() * sizeof(double));
}

void hc_he2pc_tables(double *freq, double *cdf) {
  std::vector<double> f, c;
  host::build_he2pc_table(f, c);
  memcpy(freq, f.data(), f.size() * sizeof(double));
  memcpy(cdf, c.data(), c.size() * sizeof(double));
}

/* the voxel walk on explicit packets, serial, full 16-term accumulation */
void hc_march_packets(const double *anchor, const double *sides, const int32_t *ncell,
                      const int32_t *periodic, const double *cell_n, const double *cell_xH,
                      const double *cell_xHe, int64_t np, const double *pos, const double *dir,
                      const double *sigma, const double *sigma_He_corr, const double *nu,
                      const double *weight, const double *tau, double *J, double *heat,
                      double *final_pos, int64_t *final_cell, int32_t *nsteps, int32_t max_trace,
                      int64_t *trace) {
  GridGeom g;
  for (int d = 0; d < 3; ++d) {
    g.anchor[d] = anchor[d];
    g.sides[d] = sides[d];
    g.ncell[d] = ncell[d];
    g.periodic[d] = periodic[d];
    g.cellside[d] = sides[d] / ncell[d];
    g.inv_cellside[d] = 1. / g.cellside[d];
  }
  g.ncells = (int64_t)ncell[0] * ncell[1] * ncell[2];
  const int64_t nc = g.ncells;
  const double nu_H = (13.6 * ELECTRONVOLT) * (1. / PLANCK);
  const double nu_He = (24.6 * ELECTRONVOLT) * (1. / PLANCK);
  for (int64_t p = 0; p < np; ++p) {
    MarchState s;
    s.px = pos[3 * p]; s.py = pos[3 * p + 1]; s.pz = pos[3 * p + 2];
    s.dx = dir[3 * p]; s.dy = dir[3 * p + 1]; s.dz = dir[3 * p + 2];
    s.ix_ = 1. / s.dx; s.iy_ = 1. / s.dy; s.iz_ = 1. / s.dz;
    s.tau = tau[p];
    const double *sg = sigma + p * NUM_IONS;
    march_locate(g, s);
    int32_t ns = 0;
    bool inside;
    while ((inside = march_inside(g, s)) && s.tau > 0.) {
      const int64_t cell = long_index(g, s.ix, s.iy, s.iz);
      s.last_cell = cell;
      const double ds = march_step(g, s, cell_n[cell], cell_xH[cell], cell_xHe[cell], sg[0],
                                   sigma_He_corr[p]);
      if (cell_n[cell] > 0.) {
        const double dsw = ds * weight[p];
        for (int k = 0; k < NUM_IONS; ++k) J[k * nc + cell] += dsw * sg[k];
        heat[cell] += dsw * sg[ION_H_n] * (nu[p] - nu_H);
        heat[nc + cell] += dsw * sg[ION_He_n] * (nu[p] - nu_He);
        if (trace && ns < max_trace) trace[p * max_trace + ns] = cell;
        ++ns;
      }
    }
    final_pos[3 * p] = s.px; final_pos[3 * p + 1] = s.py; final_pos[3 * p + 2] = s.pz;
    final_cell[p] = inside ? s.last_cell : -1;
    nsteps[p] = ns;
    if (trace)
      for (int32_t k = ns; k < max_trace; ++k) trace[p * max_trace + k] = -1;
  }
}

/* the product's shoot_packet (shoot.cuh) executed on the host with plain adds: logic check of
 * emission / walk / re-emission for the CPU tier.  cells [nc][4] = n, xH, xHe, T.
 * iparams: n_sources, spectrum_kind, xs_kind, reemission_kind, acc_mode(0 full,1 H-only), ranlux(0/1)
 * dparams: spectrum param, A_He, fixed reemission probability, fixed reemission frequency
 * acc: ACC_COUNTERS + nc*NACC doubles, accumulated into */
struct HostAdder {
  void operator()(double *a, double v) const { *a += v; }
};

void hc_shoot(const double *anchor, const double *sides, const int32_t *ncell,
              const int32_t *periodic, const double *cells, const int32_t *iparams,
              const double *dparams, const double *src_pos, const double *src_weights,
              const double *xs_fixed, uint64_t n_packets, uint64_t packet_offset, uint64_t seed,
              uint32_t iteration, double *acc) {
  ShootParams P;
  GridGeom &g = P.geom;
  for (int d = 0; d < 3; ++d) {
    g.anchor[d] = anchor[d];
    g.sides[d] = sides[d];
    g.ncell[d] = ncell[d];
    g.periodic[d] = periodic[d];
    g.cellside[d] = sides[d] / ncell[d];
    g.inv_cellside[d] = 1. / g.cellside[d];
  }
  g.cell_volume = g.cellside[0] * g.cellside[1] * g.cellside[2];
  g.ncells = (int64_t)ncell[0] * ncell[1] * ncell[2];
  const int64_t nc = g.ncells;
  std::vector<CellOpacity> cell_rec(nc);
  for (int64_t i = 0; i < nc; ++i) {
    cell_rec[i].n = cells[4 * i];
    cell_rec[i].xH = cells[4 * i + 1];
    cell_rec[i].xHe = cells[4 * i + 2];
    cell_rec[i].T = cells[4 * i + 3];
  }
  SourceModel &m = P.src;
  memset(&m, 0, sizeof(m));
  m.n_sources = iparams[0];
  std::vector<double> cum(m.n_sources);
  for (int i = 0; i < m.n_sources; ++i) cum[i] = (i ? cum[i - 1] : 0.) + src_weights[i];
  cum[m.n_sources - 1] = 1.;
  m.src_pos = src_pos;
  m.src_cum = cum.data();
  m.discrete_weight = 1.;
  m.spectrum.kind = iparams[1];
  std::vector<double> planck, hf, ht, hc, hef, het, hec, tf, tc;
  if (m.spectrum.kind == SPECTRUM_PLANCK) {
    host::build_planck_table(dparams[0], planck);
    m.spectrum.planck = planck.data();
  } else {
    m.spectrum.mono_frequency = dparams[0];
  }
  m.xs_kind = iparams[2];
  for (int k = 0; k < NUM_IONS; ++k) m.xs_fixed[k] = xs_fixed ? xs_fixed[k] : 0.;
  m.A_He = dparams[1];
  m.A_He_reemit = dparams[1];
  for (int k = 0; k < NUM_IONS; ++k) m.fold[k] = 1.; /* IonizationSimulation conventions */
  m.reemission_kind = iparams[3];
  m.fixed_reemission_probability = dparams[2];
  m.fixed_reemission_frequency = dparams[3];
  std::vector<double> prob;
  if (m.reemission_kind == REEMISSION_PHYSICAL) {
    const SourceModel mm = m;
    auto sig = [mm](int ion) {
      return [mm, ion](double nu) { return mm.xs_kind == XS_VERNER ? verner_cross_section(ion, nu) : mm.xs_fixed[ion]; };
    };
    host::build_lyc_table(0, sig(ION_H_n), hf, ht, hc);
    host::build_lyc_table(1, sig(ION_He_n), hef, het, hec);
    host::build_he2pc_table(tf, tc);
    m.hlyc_freq = hf.data(); m.hlyc_temp = ht.data(); m.hlyc_cdf = hc.data();
    m.helyc_freq = hef.data(); m.helyc_temp = het.data(); m.helyc_cdf = hec.data();
    m.he2pc_freq = tf.data(); m.he2pc_cdf = tc.data();
    prob.resize(nc * NUM_REEMIT);
    for (int64_t i = 0; i < nc; ++i) reemission_probabilities(cell_rec[i].T, &prob[i * NUM_REEMIT]);
  }
  P.cells = cell_rec.data();
  P.cells_h = nullptr;
  P.reemit_prob = prob.data();
  P.acc = acc;
  P.honly_cell_stride = 2;
  P.honly_term_stride = 1;
  P.honly_offset = 0;
  P.hot_acc = nullptr;
  P.src_cell = nullptr;
  P.hot_replicas = 0;
  P.nu_H = (13.6 * ELECTRONVOLT) * (1. / PLANCK);
  P.nu_He = (24.6 * ELECTRONVOLT) * (1. / PLANCK);
  P.seed = seed;
  P.iteration = iteration;
  P.packet_offset = packet_offset;
  P.n_packets = n_packets;
  ShootCounters cnt;
  const HostAdder add;
  if (iparams[5] != 0) {
    /* iparams[5]: all packets draw from ONE RANLUX stream seeded with `seed`, in order — what a
     * single-threaded IonizationPhotonShootJob does (IonizationPhotonShootJob.hpp:117-146) */
    RanluxRng rng((int)seed);
    for (uint64_t i = 0; i < n_packets; ++i) shoot_packet_from<ACC_FULL>(P, rng, add, cnt);
  } else if (iparams[4] == ACC_HONLY) {
    for (uint64_t i = 0; i < n_packets; ++i) shoot_packet<ACC_HONLY>(P, i, add, cnt);
  } else {
    for (uint64_t i = 0; i < n_packets; ++i) shoot_packet<ACC_FULL>(P, i, add, cnt);
  }
  acc[0] += cnt.w_tot;
  for (int t = 0; t < NUM_PACKET_TYPES; ++t) acc[1 + t] += cnt.w_type[t];
  acc[5] += cnt.n_steps;
  acc[6] += cnt.n_emit;
  acc[8] += cnt.tau_sum;
}

/*
 * A whole simulation on the reference's random stream: the loop body of IonizationSimulation::run
 * (IonizationSimulation.cpp:359-643) for n_iter iterations — reset, re-emission probabilities, shoot of
 * n_packets packets from ONE RANLUX generator that lives across the iterations (as the single job of a
 * single-threaded reference run does), state update of every cell (ionization-only while loop <= 3, the
 * temperature solve afterwards) — composed from the product's shared physics headers exactly as the
 * kernels compose them (shoot_packet_from; update_state_kernel's per-cell body).  One discrete source,
 * Planck spectrum, Verner cross sections and rates, Physical re-emission.  cells: [ncell][4] = n, xH, xHe, T
 * (in/out); xmetal: [ncell][12] (in/out).
 */
void hc_simulation(const double *anchor, const double *sides, const int32_t *ncell, const double *src_pos,
                   double luminosity, double planck_temperature, const double *abund, const double *tparams,
                   int do_temperature, uint32_t n_iter, uint64_t n_packets, int seed, double *cells, double *xmetal) {
  ShootParams P;
  GridGeom &g = P.geom;
  memset(&g, 0, sizeof(g));
  for (int d = 0; d < 3; ++d) {
    g.anchor[d] = anchor[d];
    g.sides[d] = sides[d];
    g.ncell[d] = ncell[d];
    g.cellside[d] = sides[d] / ncell[d];
    g.inv_cellside[d] = 1. / g.cellside[d];
  }
  g.cell_volume = g.cellside[0] * g.cellside[1] * g.cellside[2];
  g.ncells = (int64_t)ncell[0] * ncell[1] * ncell[2];
  const int64_t nc = g.ncells;
  std::vector<CellOpacity> cell_rec(nc);
  SourceModel &m = P.src;
  memset(&m, 0, sizeof(m));
  m.n_sources = 1;
  const double cum = 1.;
  m.src_pos = src_pos;
  m.src_cum = &cum;
  m.discrete_weight = 1.;
  std::vector<double> planck, hf, ht, hc, hef, het, hec, tf, tc;
  host::build_planck_table(planck_temperature, planck);
  m.spectrum.kind = SPECTRUM_PLANCK;
  m.spectrum.planck = planck.data();
  m.xs_kind = XS_VERNER;
  m.A_He = abund[EL_He];
  m.A_He_reemit = abund[EL_He];
  for (int k = 0; k < NUM_IONS; ++k) m.fold[k] = 1.; /* IonizationSimulation conventions */
  m.reemission_kind = REEMISSION_PHYSICAL;
  host::build_lyc_table(0, [](double nu) { return verner_cross_section(ION_H_n, nu); }, hf, ht, hc);
  host::build_lyc_table(1, [](double nu) { return verner_cross_section(ION_He_n, nu); }, hef, het, hec);
  host::build_he2pc_table(tf, tc);
  m.hlyc_freq = hf.data(); m.hlyc_temp = ht.data(); m.hlyc_cdf = hc.data();
  m.helyc_freq = hef.data(); m.helyc_temp = het.data(); m.helyc_cdf = hec.data();
  m.he2pc_freq = tf.data(); m.he2pc_cdf = tc.data();
  std::vector<double> prob(nc * NUM_REEMIT), acc(ACC_COUNTERS + nc * NUM_ACC);
  P.cells = cell_rec.data();
  P.cells_h = nullptr;
  P.reemit_prob = prob.data();
  P.acc = acc.data();
  P.honly_cell_stride = 2; P.honly_term_stride = 1; P.honly_offset = 0;
  P.hot_acc = nullptr; P.src_cell = nullptr; P.hot_replicas = 0;
  P.nu_H = (13.6 * ELECTRONVOLT) * (1. / PLANCK);
  P.nu_He = (24.6 * ELECTRONVOLT) * (1. / PLANCK);
  P.seed = 0; P.iteration = 0; P.packet_offset = 0; P.n_packets = n_packets;
  const RecombinationModel rr = make_rr(RR_VERNER, nullptr);
  TemperatureParams tp;
  tp.do_temperature = do_temperature;
  tp.min_iterations = 3;
  tp.pahfac = tparams[0]; tp.crfac = tparams[1]; tp.crlim = tparams[2]; tp.crscale = tparams[3];
  tp.min_ionized_T = tparams[4]; tp.epsilon = tparams[5]; tp.max_iterations = (uint32_t)tparams[6];
  RanluxRng rng(seed);
  const HostAdder add;
  for (uint32_t loop = 0; loop < n_iter; ++loop) {
    for (int64_t i = 0; i < nc; ++i) {
      cell_rec[i].n = cells[4 * i]; cell_rec[i].xH = cells[4 * i + 1];
      cell_rec[i].xHe = cells[4 * i + 2]; cell_rec[i].T = cells[4 * i + 3];
      reemission_probabilities(cell_rec[i].T, &prob[i * NUM_REEMIT]);
    }
    std::fill(acc.begin(), acc.end(), 0.);
    ShootCounters cnt;
    for (uint64_t k = 0; k < n_packets; ++k) shoot_packet_from<ACC_FULL>(P, rng, add, cnt);
    /* update_state_kernel, cell by cell (kernels.cuh) */
    const double jfac0 = luminosity / cnt.w_tot, hfac0 = jfac0 * PLANCK;
    const double jfac = jfac0 / g.cell_volume, hfac = hfac0 / g.cell_volume;
    const bool solve_T = do_temperature && loop > tp.min_iterations;
    for (int64_t i = 0; i < nc; ++i) {
      const double *a = acc.data() + ACC_COUNTERS + i * NUM_ACC;
      double J[NUM_IONS], heat[NUM_HEAT];
      for (int k = 0; k < NUM_IONS; ++k) J[k] = a[acc_slot(k)];
      heat[0] = a[acc_slot(NUM_IONS)];
      heat[1] = a[acc_slot(NUM_IONS + 1)];
      CellState out;
      if (solve_T) {
        double xprev[NUM_IONS];
        xprev[0] = cells[4 * i + 1];
        xprev[1] = cells[4 * i + 2];
        for (int k = 0; k < 12; ++k) xprev[2 + k] = xmetal[i * 12 + k];
        const int32_t iz = (int32_t)(i % g.ncell[2]);
        const double midz = (g.anchor[2] + g.cellside[2] * iz) + 0.5 * g.cellside[2];
        cell_temperature(jfac, hfac, J, heat, cells[4 * i], cells[4 * i + 3], -1., midz, abund, rr, tp, xprev, out);
      } else {
        cell_ionization_state(jfac, hfac, J, heat, cells[4 * i], cells[4 * i + 3], abund, rr, out);
      }
      cells[4 * i + 1] = out.x[ION_H_n];
      cells[4 * i + 2] = out.x[ION_He_n];
      cells[4 * i + 3] = out.T;
      for (int k = 0; k < 12; ++k) xmetal[i * 12 + k] = out.x[2 + k];
    }
  }
}

} /* extern "C" */
