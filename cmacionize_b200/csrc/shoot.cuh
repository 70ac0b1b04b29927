/*
 * shoot.cuh — life of one photon packet: emit -> tau -> voxel walk with
 * accumulation -> (re-emit -> tau -> walk)* .
 *
 * Behavioural contract = IonizationPhotonShootJob::execute
 * (/root/reference/src/IonizationPhotonShootJob.hpp:117-146) with
 *   PhotonSource::get_random_photon   src/PhotonSource.cpp:208-249
 *   CartesianDensityGrid::interact    src/CartesianDensityGrid.cpp:375-452
 *   DensityGrid::update_integrals     src/DensityGrid.hpp:150-197
 *   PhotonSource::reemit              src/PhotonSource.cpp:272-308
 *
 * Written once as a __host__ __device__ function over an `Adder` policy so that
 * the kernels (atomic RED adds) and the CPU-tier logic check of the tests
 * (tests/hostcheck, plain +=) execute the same statements.
 *
 * Accumulator layouts: after ACC_COUNTERS leading counter doubles,
 *   ACC_FULL   acc[cell][16] = J[14], heat_H, heat_He in the order of acc_slot()  (128 B = one L2 line per cell)
 *   ACC_HONLY  J_H, heat_H per cell (used when only sigma_H != 0), interleaved or as two planes
 *              (ShootParams::honly_*_stride): a monochromatic source at the threshold adds no
 *              heat, so with planes only the 8 B/cell J plane is ever touched — half the
 *              footprint competing for L2 on HBM-resident grids
 */
#pragma once
#include "cmib_common.cuh"
#include "march.cuh"
#include "rng.cuh"
#include "source.cuh"

namespace cmib {

enum AccMode : int { ACC_FULL = 0, ACC_HONLY = 1 };
/* leading counters: totweight, typecount[4], cell crossings, (re)emissions, accumulator adds,
 * optical depth traversed, 7 spare.  16 doubles = 128 B so that the per-cell records that follow
 * are aligned to L2 lines (with 8 counters every cell's 16 accumulators straddled two lines) */
constexpr int ACC_COUNTERS = 16;

/*
 * Position of term k (0..13 = J of ion k, 14 = heat_H, 15 = heat_He) inside a cell's 16-double
 * record of the full layout.  Terms that are added together share a 32-byte sector: a photon
 * between 13.6 and 21.6 eV adds J_H, heat_H, J_O0, J_N0 — one sector instead of three — which is
 * what an HBM-resident grid pays for (every touched sector is a read-modify-write in DRAM).
 *   sector 0: J_H heat_H J_O0 J_N0        (thresholds 13.6, 13.6, 14.5 eV)
 *   sector 1: J_He heat_He J_Ne0 J_S+     (24.6, 21.6, 23.3 eV)
 *   sector 2: J_C+ J_N+ J_S++ J_O+        (24.4, 29.6, 34.8, 35.1 eV)
 *   sector 3: J_Ne+ J_N++ J_S+++ J_C++    (41.0, 47.4, 47.2, 47.9 eV)
 * packed as 4-bit fields of a 64-bit constant so that a run-time k costs a shift and a mask.
 */
constexpr uint64_t acc_slot_table() {
  const int slot_of_term[16] = {/*H*/ 0, /*He*/ 4, /*C+*/ 8, /*C++*/ 15, /*N0*/ 3, /*N+*/ 9, /*N++*/ 13, /*O0*/ 2,
                                /*O+*/ 11, /*Ne0*/ 6, /*Ne+*/ 12, /*S+*/ 7, /*S++*/ 10, /*S+++*/ 14,
                                /*heat_H*/ 1, /*heat_He*/ 5};
  uint64_t t = 0;
  for (int k = 0; k < 16; ++k) t |= (uint64_t)slot_of_term[k] << (4 * k);
  return t;
}
constexpr uint64_t ACC_SLOT_TABLE = acc_slot_table();
CMIB_HD int acc_slot(int k) { return (int)((ACC_SLOT_TABLE >> (4 * k)) & 15u); }

template <int MODE> struct AccLayout;
template <> struct AccLayout<ACC_FULL> { static constexpr int NACC = 16; static constexpr int NSIG = 14; };
template <> struct AccLayout<ACC_HONLY> { static constexpr int NACC = 2; static constexpr int NSIG = 1; };

struct ShootParams {
  GridGeom geom;
  SourceModel src;
  const CellOpacity *cells;
  const double2 *cells_h;    /* compact (n, x_H) copy for the H-only walk: 16 B instead of 32 B per gather */
  const double *reemit_prob; /* [ncell][5] (REEMISSION_PHYSICAL) */
  double *acc;               /* counters + per-cell accumulators */
  /* H-only layout: term k of cell c lives at
   * acc[ACC_COUNTERS + honly_offset + c*honly_cell_stride + k*honly_term_stride]:
   * interleaved (2, 1) while the grid is L2 resident (spreads hot cells over more sectors),
   * planar (1, ncells) when it is not (halves the footprint a threshold source touches) */
  int64_t honly_cell_stride, honly_term_stride, honly_offset;
  double nu_H, nu_He;        /* 13.6 eV, 24.6 eV in Hz (DensityGrid.hpp:219-222) */
  /* hot-cell replication (wavefront path): the 3x3x3 cells around every source receive the first
   * crossings of ALL its packets; their accumulators are replicated hot_replicas times
   * (hot_acc[replica][source][27][16]) and folded into acc at the end of the shoot */
  double *hot_acc;
  const uint32_t *src_cell; /* packed cell indices of the sources: ix | iy << 10 | iz << 20 */
  int hot_replicas;         /* 0: off */
  uint64_t seed;
  uint32_t iteration;
  uint64_t packet_offset;
  uint64_t n_packets;
};

/* per-thread partial sums of IonizationPhotonShootJob's counters + roofline diagnostics */
struct ShootCounters {
  double w_tot = 0.;
  double w_type[NUM_PACKET_TYPES] = {0., 0., 0., 0.};
  uint32_t n_steps = 0, n_emit = 0; /* cell crossings, (re)emissions */
  uint32_t n_red = 0;               /* accumulator terms added (wavefront path only) */
  double tau_sum = 0.;              /* optical depth traversed: sum of tau_cell over all crossings */
};

/* update_integrals (DensityGrid.hpp:150-197): zero increments are skipped, which
 * is exact (x + 0.0 == x) and removes most of the 16 RMWs for soft photons */
/* address of accumulator term k (FULL: 0..13 J, 14 heat_H, 15 heat_He; HONLY: 0 J_H, 1 heat_H) */
template <int MODE>
CMIB_HD double *acc_term(const ShootParams &P, int64_t cell, int k) {
  if (MODE == ACC_HONLY)
    return P.acc + ACC_COUNTERS + P.honly_offset + cell * P.honly_cell_stride + (int64_t)k * P.honly_term_stride;
  return P.acc + ACC_COUNTERS + cell * AccLayout<MODE>::NACC + acc_slot(k);
}

template <int MODE, class Adder>
CMIB_HD void accumulate(const Adder &add, const ShootParams &P, int64_t cell, double ds, double weight,
                        const double *sigma, double dnu_H, double dnu_He) {
  const double dsw = ds * weight;
  double *a = P.acc + ACC_COUNTERS + cell * AccLayout<MODE>::NACC;
  if (MODE == ACC_HONLY) {
    const double dJ = dsw * sigma[0];
    add(acc_term<MODE>(P, cell, 0), dJ);
    const double dh = dJ * dnu_H;
    if (dh != 0.) add(acc_term<MODE>(P, cell, 1), dh);
  } else {
    const double dJH = dsw * sigma[ION_H_n];
    const double dJHe = dsw * sigma[ION_He_n];
#pragma unroll
    for (int ion = 0; ion < NUM_IONS; ++ion) {
      const double dJ = dsw * sigma[ion];
      if (dJ != 0.) add(a + acc_slot(ion), dJ);
    }
    const double dhH = dJH * dnu_H;
    if (dhH != 0.) add(a + acc_slot(NUM_IONS + HEAT_H), dhH);
    const double dhHe = dJHe * dnu_He;
    if (dhHe != 0.) add(a + acc_slot(NUM_IONS + HEAT_He), dhHe);
  }
}

CMIB_HD CellOpacity load_cell(const CellOpacity *cells, int64_t cell) {
#if defined(__CUDA_ARCH__)
  const double2 *cp = reinterpret_cast<const double2 *>(cells + cell);
  const double2 r0 = __ldg(cp), r1 = __ldg(cp + 1);
  CellOpacity c;
  c.n = r0.x; c.xH = r0.y; c.xHe = r1.x; c.T = r1.y;
  return c;
#else
  return cells[cell];
#endif
}

/* The walks of one packet until it leaves the box or is absorbed for good (IonizationPhotonShootJob.hpp:135-146:
 * interact, reemit, repeat).  `s` holds position and direction, sigma / sigma_He_corr / nu / type / weight the packet.
 * reemit_first = false: the packet has just been emitted and walks first; true: it was absorbed in cell s.last_cell
 * whose record is `c` (the tail of the wavefront pipeline resumes packets there) and the re-emission decision comes
 * first. */
template <int MODE, class Adder, class Rng>
CMIB_HD void packet_walks(const ShootParams &P, Rng &rng, const Adder &add, ShootCounters &cnt, MarchState &s,
                          double *sigma, double &sigma_He_corr, double &nu, int &type, double weight, bool reemit_first,
                          const CellOpacity &c_absorbed) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  const GridGeom &g = P.geom;
  const SourceModel &m = P.src;
  CellOpacity c = c_absorbed;
  bool alive = true;
  bool walk = !reemit_first;
  while (alive) {
    if (walk) {
      ++cnt.n_emit;
      s.ix_ = 1. / s.dx;
      s.iy_ = 1. / s.dy;
      s.iz_ = 1. / s.dz;
      s.tau = -log(rng_uniform(rng));
      const double tau0 = s.tau;
      march_locate(g, s);
      const double dnu_H = nu - P.nu_H;
      const double dnu_He = nu - P.nu_He;
      c.n = c.xH = c.xHe = c.T = 0.;
      bool inside;
      while ((inside = march_inside(g, s)) && s.tau > 0.) {
        const int64_t cell = long_index(g, s.ix, s.iy, s.iz);
        s.last_cell = cell;
        c = load_cell(P.cells, cell);
        const double ds = march_step(g, s, c.n, c.xH, c.xHe, sigma[0], sigma_He_corr);
        if (c.n > 0.) accumulate<MODE>(add, P, cell, ds, weight, sigma, dnu_H, dnu_He);
        ++cnt.n_steps;
      }
      /* optical depth traversed by this walk: all of it when absorbed, the used part when it left */
      cnt.tau_sum += tau0 - ((s.tau > 0.) ? s.tau : 0.);
      if (!inside) break; /* left the box: keeps its last type */
    }
    walk = true;
    /* --- PhotonSource::reemit --- */
    double new_nu = 0.;
    if (m.reemission_kind == REEMISSION_PHYSICAL) {
      double p[NUM_REEMIT];
#pragma unroll
      for (int k = 0; k < NUM_REEMIT; ++k) p[k] = P.reemit_prob[s.last_cell * NUM_REEMIT + k];
      /* ACC_HONLY is only selected when sigma_He == 0 */
      const double sHe = (NSIG > 1) ? sigma[(NSIG > 1) ? ION_He_n : 0] : 0.;
      new_nu = physical_reemit(m, sigma[0], sHe, c.xH, c.xHe, c.T, p, rng, type);
    } else if (m.reemission_kind == REEMISSION_FIXED) {
      const double u = rng_uniform(rng);
      if (u < m.fixed_reemission_probability) {
        type = PACKET_DIFFUSE_HI;
        new_nu = m.fixed_reemission_frequency;
      } else {
        type = PACKET_ABSORBED;
      }
    } else {
      type = PACKET_ABSORBED;
    }
    if (new_nu == 0.) {
      alive = false;
    } else {
      nu = new_nu;
      random_direction(rng, s.dx, s.dy, s.dz);
      packet_cross_sections<NSIG>(m, nu, sigma, sigma_He_corr);
    }
  }
  cnt.w_tot += weight;
#pragma unroll
  for (int t = 0; t < NUM_PACKET_TYPES; ++t) cnt.w_type[t] += (t == type) ? weight : 0.;
}

/* the life of one packet drawing from `rng`: the kernels pass the packet's own Philox stream, the CPU-tier
 * test passes the reference's RANLUX stream shared by all packets of a thread, as IonizationPhotonShootJob does */
template <int MODE, class Adder, class Rng>
CMIB_HD void shoot_packet_from(const ShootParams &P, Rng &rng, const Adder &add, ShootCounters &cnt) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  const SourceModel &m = P.src;
  MarchState s;
  double sigma[NSIG];
  double sigma_He_corr;
  double nu;
  int type = PACKET_PRIMARY;
  /* --- PhotonSource::get_random_photon --- */
  int isrc;
  emit_primary(m, P.geom, rng, s.px, s.py, s.pz, s.dx, s.dy, s.dz, nu, isrc);
  const double weight = (isrc >= 0) ? m.discrete_weight : m.continuous_weight;
  packet_cross_sections<NSIG>(m, nu, sigma, sigma_He_corr);
  const CellOpacity none = {0., 0., 0., 0.};
  packet_walks<MODE>(P, rng, add, cnt, s, sigma, sigma_He_corr, nu, type, weight, false, none);
}

/* packet `i` of this call (global id P.packet_offset + i) */
template <int MODE, class Adder>
CMIB_HD void shoot_packet(const ShootParams &P, uint64_t i, const Adder &add, ShootCounters &cnt) {
  PacketRng rng;
  rng_init(rng, P.seed, P.iteration, P.packet_offset + i);
  shoot_packet_from<MODE>(P, rng, add, cnt);
}

} // namespace cmib
