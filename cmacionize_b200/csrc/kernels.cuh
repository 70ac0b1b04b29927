/*
 * kernels.cuh — kernels of the photoionization hot path other than the shoot pipeline
 * (wavefront.cuh).
 *
 *   shoot_kernel          one thread per packet (shoot_packet of shoot.cuh): A/B check of the pipeline
 *   march_packets_kernel  test hook: the walk on caller-supplied packets
 *   reemission_probabilities_kernel                          per-cell pass before a shoot
 *   update_state_kernel        ionization-only update, one thread per cell
 *   update_temperature_kernel  temperature solve: persistent warps, three lanes per cell
 *   pack / unpack kernels      host SoA <-> device layouts
 *   eval_* kernels             element-wise probes of the physics for the parity tests
 *
 * Accumulator layouts: after ACC_COUNTERS leading counter doubles,
 *   ACC_FULL   acc[cell][16] = J[14], heat_H, heat_He in acc_slot() order (one 128-B L2 line per cell)
 *   ACC_HONLY  J_H, heat_H per cell: padded to a line, interleaved, or as planes (shoot.cuh)
 */
#pragma once
#include <cuda_runtime.h>

#include "cmib_common.cuh"
#include "march.cuh"
#include "rng.cuh"
#include "shoot.cuh"
#include "source.cuh"
#include "state.cuh"

namespace cmib {

/* AccMode / AccLayout / ShootParams / accumulate / shoot_packet: shoot.cuh */

/* fire-and-forget FP64 add: compiles to RED.E.ADD.F64 (no return value) */
struct DeviceAdder {
  __device__ __forceinline__ void operator()(double *addr, double v) const { atomicAdd(addr, v); }
};

/* ------------------------------------------------------------------------- */
/* shoot: one thread follows one packet at a time (grid-stride over packet ids) */
/* ------------------------------------------------------------------------- */
template <int MODE>
__global__ void __launch_bounds__(256)
shoot_kernel(const __grid_constant__ ShootParams P) {
  ShootCounters cnt;
  const DeviceAdder add;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < P.n_packets; i += stride)
    shoot_packet<MODE>(P, i, add, cnt);
  const double w_tot = cnt.w_tot;
  const double *w_type = cnt.w_type;
  const uint32_t n_steps = cnt.n_steps, n_emit = cnt.n_emit;

  /* IonizationPhotonShootJobMarket::update_counters: block reduce, one RED per block */
  __shared__ double red[9][8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[9] = {w_tot, w_type[0], w_type[1], w_type[2], w_type[3], (double)n_steps, (double)n_emit, 0., cnt.tau_sum};
#pragma unroll
  for (int k = 0; k < 9; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double sum = 0.;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) sum += red[threadIdx.x][w];
    if (sum != 0.) atomicAdd(P.acc + threadIdx.x, sum);
  }
}

/* ------------------------------------------------------------------------- */
/* test hook: CartesianDensityGrid::interact on explicit packets              */
/* ------------------------------------------------------------------------- */
struct MarchPacketsParams {
  GridGeom geom;
  const CellOpacity *cells;
  double *acc;
  double nu_H, nu_He;
  int64_t np;
  const double *pos, *dir, *sigma, *sigma_He_corr, *nu, *weight, *tau;
  double *final_pos;
  int64_t *final_cell;
  int32_t *nsteps;
  int32_t max_trace;
  int64_t *trace;
};

__global__ void __launch_bounds__(128)
march_packets_kernel(const __grid_constant__ MarchPacketsParams P) {
  const GridGeom &g = P.geom;
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P.np) return;
  MarchState s;
  s.px = P.pos[3 * p]; s.py = P.pos[3 * p + 1]; s.pz = P.pos[3 * p + 2];
  s.dx = P.dir[3 * p]; s.dy = P.dir[3 * p + 1]; s.dz = P.dir[3 * p + 2];
  s.ix_ = 1. / s.dx; s.iy_ = 1. / s.dy; s.iz_ = 1. / s.dz;
  s.tau = P.tau[p];
  double sigma[NUM_IONS];
#pragma unroll
  for (int k = 0; k < NUM_IONS; ++k) sigma[k] = P.sigma[p * NUM_IONS + k];
  const double sHe = P.sigma_He_corr[p];
  const double nu = P.nu[p];
  const double w = P.weight[p];
  march_locate(g, s);
  int32_t nsteps = 0;
  bool inside;
  while ((inside = march_inside(g, s)) && s.tau > 0.) {
    const int64_t cell = long_index(g, s.ix, s.iy, s.iz);
    s.last_cell = cell;
    const CellOpacity c = P.cells[cell];
    const double ds = march_step(g, s, c.n, c.xH, c.xHe, sigma[0], sHe);
    if (c.n > 0.) {
      ShootParams sp;
      sp.acc = P.acc;
      accumulate<ACC_FULL>(DeviceAdder(), sp, cell, ds, w, sigma, nu - P.nu_H, nu - P.nu_He);
      if (P.trace && nsteps < P.max_trace) P.trace[p * P.max_trace + nsteps] = cell;
      ++nsteps;
    }
  }
  /* interact() re-evaluates is_inside after the loop (:447): idempotent here */
  P.final_pos[3 * p] = s.px; P.final_pos[3 * p + 1] = s.py; P.final_pos[3 * p + 2] = s.pz;
  P.final_cell[p] = inside ? s.last_cell : -1;
  P.nsteps[p] = nsteps;
  if (P.trace)
    for (int32_t k = nsteps; k < P.max_trace; ++k) P.trace[p * P.max_trace + k] = -1;
}

/* DensityGrid::integrate_optical_depth for caller-supplied packets, one thread per packet */
__global__ void __launch_bounds__(128)
integrate_optical_depth_kernel(GridGeom g, const CellOpacity *cells, int64_t np, const double *pos, const double *dir,
                               const double *sigma_H, const double *sigma_He_corr, double *tau) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= np) return;
  MarchState s;
  s.px = pos[3 * p]; s.py = pos[3 * p + 1]; s.pz = pos[3 * p + 2];
  s.dx = dir[3 * p]; s.dy = dir[3 * p + 1]; s.dz = dir[3 * p + 2];
  const int64_t limit = 1ll << 22; /* only a ray that never leaves a periodic box gets here */
  tau[p] = integrate_optical_depth(g, s, sigma_H[p], sigma_He_corr[p], [cells](int64_t c) { return cells[c]; }, limit);
}

/* ------------------------------------------------------------------------- */
/* per-cell passes                                                            */
/* ------------------------------------------------------------------------- */
__global__ void reemission_probabilities_kernel(int64_t ncell, const CellOpacity *cells,
                                                double *prob) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  double p[NUM_REEMIT];
  reemission_probabilities(cells[i].T, p);
#pragma unroll
  for (int k = 0; k < NUM_REEMIT; ++k) prob[i * NUM_REEMIT + k] = p[k];
}

struct UpdateParams {
  GridGeom geom;
  CellOpacity *cells;
  double2 *cells_h;      /* compact (n, x_H) copy, kept in step with cells */
  double *xmetal;        /* [ncell][12] */
  double *heat_norm;     /* [ncell][2] */
  const double *cr_factor; /* [ncell] or NULL */
  const double *acc;
  int64_t honly_cell_stride, honly_term_stride, honly_offset; /* H-only accumulator layout (shoot.cuh) */
  double luminosity;
  double totweight;      /* <= 0: read acc[0] */
  double abund[NUM_ELEMENTS];
  RecombinationModel rr;
  TemperatureParams tp;
  int solve_temperature;
  int64_t cell_begin, cell_end; /* the cell block this launch updates (MPICommunicator::distribute_block) */
  /* work items j = 0 .. n_work - 1 of this launch.  own_size <= 1: cell = cell_begin + j (a contiguous block).
   * own_size > 1: the rank's share of a multi-GPU update — the grid is cut into chunks of OWN_CHUNK cells and
   * chunk c belongs to rank c % own_size, so that the expensive cells (the temperature solve inside the ionised
   * region) are spread over all ranks; cell = ((j / OWN_CHUNK) * own_size + own_rank) * OWN_CHUNK + j % OWN_CHUNK,
   * items whose cell lies behind cell_end (the tail of the last chunk) are skipped */
  int64_t n_work;
  int32_t own_rank, own_size;
  /* TaskBasedIonizationSimulation.cpp:932-951: the packets carried abundance-weighted cross sections; the mean
   * intensities of every ion but H0 are divided by the abundance of the ion's element (where it is positive), the
   * helium heating term by the helium abundance, before the state is computed */
  int fold_abundances;
  int lc_wide_max_pairs; /* update_temperature_kernel: line cooling warp-wide when a warp holds at most this many
                          * (cell, temperature) pairs (LC_WIDE_MAX_PAIRS; 0 = never, CMIB_LC_WIDE for A/B runs) */
};

/* element of ion k >= 1 in the order of `Element` (ElementNames.hpp get_element) */
CMIB_HD int ion_element(int ion) {
  return (ion == ION_He_n) ? EL_He : (ion <= ION_C_p2) ? EL_C : (ion <= ION_N_p2) ? EL_N : (ion <= ION_O_p1) ? EL_O
         : (ion <= ION_Ne_p1) ? EL_Ne : EL_S;
}
CMIB_HD void unfold_abundances(const double *abund, double *J, double *heat) {
#pragma unroll
  for (int ion = 1; ion < NUM_IONS; ++ion) {
    const double a = abund[ion_element(ion)];
    if (a > 0.) J[ion] = J[ion] / a;
  }
  if (abund[EL_He] > 0.) heat[1] = heat[1] / abund[EL_He];
}

constexpr int64_t OWN_CHUNK = 1024;
CMIB_HD int64_t owned_cell(int64_t j, int32_t rank, int32_t size) {
  return ((j / OWN_CHUNK) * size + rank) * OWN_CHUNK + (j % OWN_CHUNK);
}
CMIB_D int64_t work_cell(const UpdateParams &P, int64_t j) {
  return (P.own_size <= 1) ? P.cell_begin + j : owned_cell(j, P.own_rank, P.own_size);
}

template <int MODE>
__global__ void __launch_bounds__(128)
update_state_kernel(const __grid_constant__ UpdateParams P) {
  const int64_t jw = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (jw >= P.n_work) return;
  const int64_t i = work_cell(P, jw);
  if (i >= P.cell_end) return;
  const double totweight = (P.totweight > 0.) ? P.totweight : P.acc[0];
  /* jfac = L / W, hfac = jfac * h, both divided by the cell volume per cell
   * (IonizationStateCalculator.cpp:519-521, .hpp:135-139) */
  const double jfac0 = P.luminosity / totweight;
  const double hfac0 = jfac0 * PLANCK;
  const double jfac = jfac0 / P.geom.cell_volume;
  const double hfac = hfac0 / P.geom.cell_volume;
  double J[NUM_IONS], heat[NUM_HEAT];
  const double *a = P.acc + ACC_COUNTERS + i * AccLayout<MODE>::NACC;
  if (MODE == ACC_HONLY) {
#pragma unroll
    for (int k = 0; k < NUM_IONS; ++k) J[k] = 0.;
    J[0] = P.acc[ACC_COUNTERS + P.honly_offset + i * P.honly_cell_stride];
    heat[0] = P.acc[ACC_COUNTERS + P.honly_offset + i * P.honly_cell_stride + P.honly_term_stride];
    heat[1] = 0.;
  } else {
#pragma unroll
    for (int k = 0; k < NUM_IONS; ++k) J[k] = a[acc_slot(k)];
    heat[0] = a[acc_slot(NUM_IONS)];
    heat[1] = a[acc_slot(NUM_IONS + 1)];
  }
  if (P.fold_abundances) unfold_abundances(P.abund, J, heat);
  CellOpacity c = P.cells[i];
  CellState out;
  if (P.solve_temperature) {
    double xprev[NUM_IONS];
    xprev[0] = c.xH;
    xprev[1] = c.xHe;
#pragma unroll
    for (int k = 0; k < 12; ++k) xprev[2 + k] = P.xmetal[i * 12 + k];
    /* cell midpoint z (CartesianDensityGrid.hpp:85-89): anchor + cellside*iz + 0.5*cellside */
    const int32_t iz = (int32_t)(i % P.geom.ncell[2]);
    const double midz = (P.geom.anchor[2] + P.geom.cellside[2] * iz) + 0.5 * P.geom.cellside[2];
    const double crf = P.cr_factor ? P.cr_factor[i] : -1.;
    cell_temperature(jfac, hfac, J, heat, c.n, c.T, crf, midz, P.abund, P.rr, P.tp, xprev, out);
  } else {
    cell_ionization_state(jfac, hfac, J, heat, c.n, c.T, P.abund, P.rr, out);
  }
  c.xH = out.x[ION_H_n];
  c.xHe = out.x[ION_He_n];
  c.T = out.T;
  P.cells[i] = c;
  P.cells_h[i] = make_double2(c.n, c.xH);
#pragma unroll
  for (int k = 0; k < 12; ++k) P.xmetal[i * 12 + k] = out.x[2 + k];
  P.heat_norm[i * 2] = out.heat[0];
  P.heat_norm[i * 2 + 1] = out.heat[1];
}

/*
 * Temperature update: persistent warps, dynamic cell hand-out, THREE lanes per cell.
 *
 * A secant iteration of TemperatureCalculator::calculate_temperature evaluates the heating/cooling
 * balance at 1.1 T0, 0.9 T0 and T0 (TemperatureCalculator.cpp:730-760): three independent
 * evaluations of the same inputs.  A cell owns three adjacent lanes (10 cells per warp, lanes 30/31
 * idle), lane role r evaluates temperature r, the six numbers of the secant update are exchanged
 * with shuffles and all three lanes advance the same state.  Every pass is one balance evaluation
 * for every lane that has a cell; a cell that converged is stored and its lanes take the next
 * unprocessed cell from a global counter.
 *
 * Why: (1) with one thread bound to one cell (update_state_kernel) a warp runs until its slowest
 * cell has converged and cells that need no solve idle a lane for the whole time — ncu: 17 of 32
 * lanes active (profiles/r01_update_state.md); (2) a few cells at the ionisation front run the full
 * 100 iterations; with three sequential evaluations per iteration they alone kept the kernel alive
 * for ~20 ms after everything else had finished (profiles/r01_update_state.md) — their critical
 * path is three times shorter here.  The arithmetic per cell is unchanged (bitwise equal to the
 * per-cell kernel, tests/test_gpu_simulation.py).
 */
/*
 * Line cooling of the LAST cell of a warp, warp-wide.  A cell that runs many secant iterations is a chain of balance
 * evaluations, and two thirds of the instructions (one third of the latency) of an evaluation is the line cooling: ten five-level systems (10 collision strengths,
 * 10 Boltzmann factors, a 5 x 5 solve each) and three two-level ones, one after the other on the cell's lane.  While a
 * warp is full that is the right shape (every lane is busy); when only a few (cell, temperature) pairs are left the other
 * lanes idle and the chain is what the whole state update waits for (profiles/r02_exchange.md: 2.0 ms for an eighth of
 * the cells at 8 GPUs).  Here up to three pairs per round are served by ten lanes each: lane k of a group evaluates
 * five-level element k (lanes 0-2 also two-level element k) for its pair from the pair's T, n_e and abundances
 * (shuffles), the owner collects the 13 terms and adds them in the order of line_cooling().  Every term is computed by
 * the code line_cooling() runs (line_cooling_term5 / term2), only on another lane and with the table read from shared
 * memory (the lanes of a warp read different elements): the result is bit-identical.
 *   pairs: lanes that hold a pair; has: this lane does; T, ne, ab: the pair's inputs (BalanceMid). */
constexpr int LC_WIDE_MAX_PAIRS = 3; /* the three temperatures of ONE cell: a single round.  Measured (lexingtonHII20 64^3,
                                      * mean state update over 24 iterations): never 4.13 ms, 3 pairs 3.47 ms, 9 pairs 3.61 ms;
                                      * an iteration with a 100-iteration cell 9.3 -> 6.8 ms (profiles/r02_update.md) */
__device__ __noinline__ double line_cooling_wide(const double *s_tab, unsigned pairs, bool has, double T, double ne, const double *ab) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int g = lane / 10, k = lane - 10 * g; /* group 3 = lanes 30, 31: no work */
  const LcTablePtr tab = {s_tab};
  double result = 0.;
  unsigned todo = pairs;
  while (todo != 0u) {
    const int o0 = __ffs(todo) - 1;
    todo &= todo - 1u;
    const int o1 = todo ? __ffs(todo) - 1 : -1;
    if (todo) todo &= todo - 1u;
    const int o2 = todo ? __ffs(todo) - 1 : -1;
    if (todo) todo &= todo - 1u;
    const int src = (g == 0) ? o0 : ((g == 1) ? o1 : ((g == 2) ? o2 : -1));
    const int from = (src < 0) ? lane : src;
    const double Ts = __shfl_sync(full, T, from), nes = __shfl_sync(full, ne, from);
    double ab5 = 0., ab2 = 0.;
#pragma unroll
    for (int i = 0; i < LC_NUM; ++i) {
      const double v = __shfl_sync(full, ab[i], from);
      if (i == k) ab5 = v;
      if (i == LC_NUM5 + k) ab2 = v;
    }
    double t5 = 0., t2 = 0.;
    if (src >= 0 && nes != 0.) {
      const double prefactor = tab[LC_OFF_PREFACTOR] * nes / sqrt(Ts);
      const double Tinv = 1. / Ts;
      const double logT = log(Ts);
      t5 = line_cooling_term5(tab, k, prefactor, Ts, Tinv, logT, ab5);
      if (k < 3) t2 = line_cooling_term2(tab, k, prefactor, Ts, Tinv, logT, ab2);
    }
    /* owners collect: group of the owner = its rank in this round */
    const int mine = (lane == o0) ? 0 : ((lane == o1) ? 10 : ((lane == o2) ? 20 : -1));
    const int base = (mine < 0) ? 0 : mine;
    double cooling = 0.;
#pragma unroll
    for (int e = 0; e < LC_NUM5; ++e) cooling += __shfl_sync(full, t5, base + e);
#pragma unroll
    for (int i = 0; i < 3; ++i) cooling += __shfl_sync(full, t2, base + i);
    if (mine >= 0 && has) result = (ne == 0.) ? 1.e-99 : cooling;
  }
  return result;
}

template <int MODE>
__global__ void __launch_bounds__(128)
update_temperature_kernel(const __grid_constant__ UpdateParams P, unsigned long long *next_cell) {
  __shared__ double s_lc_tab[LC_TABLE_SIZE];
  for (int i = threadIdx.x; i < LC_TABLE_SIZE; i += blockDim.x) s_lc_tab[i] = CMIB_TBL(LINECOOLING)[i];
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const bool in_slot = lane < 30;
  const int role = lane % 3;            /* 0: 1.1 T0, 1: 0.9 T0, 2: T0 */
  const int slot_base = lane - role;    /* first lane of this cell's three */
  const unsigned role0_mask = 0x09249249u; /* lanes 0, 3, ..., 27 */
  const int64_t nwork = P.n_work; /* next_cell counts work items from 0 */
  const int wide_max = P.lc_wide_max_pairs;
  const double totweight = (P.totweight > 0.) ? P.totweight : P.acc[0];
  const double jfac = (P.luminosity / totweight) / P.geom.cell_volume;
  const double hfac = ((P.luminosity / totweight) * PLANCK) / P.geom.cell_volume;
  TemperatureSolve S;
  CellState out;
  double j[NUM_IONS], h[NUM_HEAT];
  double ntot = 0., midz = 0.;
  int64_t cell = -1;
  bool jH_zero = false, jHe_zero = false;
  bool has = false, exhausted = false;

  auto store = [&](int64_t i, double n_keep) {
    CellOpacity c;
    c.n = n_keep;
    c.xH = out.x[ION_H_n];
    c.xHe = out.x[ION_He_n];
    c.T = out.T;
    P.cells[i] = c;
    P.cells_h[i] = make_double2(c.n, c.xH);
#pragma unroll
    for (int k = 0; k < 12; ++k) P.xmetal[i * 12 + k] = out.x[2 + k];
    P.heat_norm[i * 2] = out.heat[0];
    P.heat_norm[i * 2 + 1] = out.heat[1];
  };

  while (true) {
    /* ---- hand cells to idle slots; cells that need no solve are finished on the spot ---- */
    for (int tries = 0; tries < 8; ++tries) {
      const unsigned idle = __ballot_sync(0xffffffffu, in_slot && !has) & role0_mask; /* one bit per idle slot */
      if (idle == 0u || exhausted) break;
      unsigned long long base = 0;
      const int leader = __ffs(idle) - 1;
      if (lane == leader) base = atomicAdd(next_cell, (unsigned long long)__popc(idle));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (base + __popc(idle) >= (unsigned long long)nwork) exhausted = true;
      const int64_t jw = (int64_t)base + __popc(idle & ((1u << slot_base) - 1u));
      const int64_t i = (jw < nwork) ? work_cell(P, jw) : P.cell_end;
      if (in_slot && !has && i < P.cell_end) {
        /* the three lanes of the slot load the same cell and begin the same solve */
        double J[NUM_IONS], heat[NUM_HEAT], xprev[NUM_IONS];
        if (MODE == ACC_HONLY) {
#pragma unroll
          for (int k = 0; k < NUM_IONS; ++k) J[k] = 0.;
          J[0] = P.acc[ACC_COUNTERS + P.honly_offset + i * P.honly_cell_stride];
          heat[0] = P.acc[ACC_COUNTERS + P.honly_offset + i * P.honly_cell_stride + P.honly_term_stride];
          heat[1] = 0.;
        } else {
          const double *a = P.acc + ACC_COUNTERS + i * AccLayout<MODE>::NACC;
#pragma unroll
          for (int k = 0; k < NUM_IONS; ++k) J[k] = a[acc_slot(k)];
          heat[0] = a[acc_slot(NUM_IONS)];
          heat[1] = a[acc_slot(NUM_IONS + 1)];
        }
        if (P.fold_abundances) unfold_abundances(P.abund, J, heat);
        const CellOpacity c = P.cells[i];
        xprev[0] = c.xH;
        xprev[1] = c.xHe;
#pragma unroll
        for (int k = 0; k < 12; ++k) xprev[2 + k] = P.xmetal[i * 12 + k];
        const double crf = P.cr_factor ? P.cr_factor[i] : -1.;
        if (temperature_solve_begin(S, jfac, hfac, J, heat, c.n, c.T, crf, P.abund, P.rr, P.tp, xprev, j, h, out)) {
          if (temperature_solve_continues(S, P.tp)) {
            /* cell midpoint z (CartesianDensityGrid.hpp:85-89): anchor + cellside*iz + 0.5*cellside */
            const int32_t iz = (int32_t)(i % P.geom.ncell[2]);
            midz = (P.geom.anchor[2] + P.geom.cellside[2] * iz) + 0.5 * P.geom.cellside[2];
            ntot = c.n;
            cell = i;
            jH_zero = (J[ION_H_n] == 0.);
            jHe_zero = (J[ION_He_n] == 0.);
            has = true;
          } else if (role == 2) {
            temperature_solve_finish(S, J, h, out);
            store(i, c.n);
          }
        } else if (role == 2) {
          store(i, c.n);
        }
      }
    }
    if (__ballot_sync(0xffffffffu, has) == 0u) {
      if (exhausted) break;
      continue;
    }
    /* ---- one balance evaluation per lane: the slot's three temperatures side by side ---- */
    double h0e = 0., he0e = 0., gain = 0., loss = 0.;
    BalanceMid mid;
    mid.T = 1.; mid.ne = 0.;
#pragma unroll
    for (int i = 0; i < LC_NUM; ++i) mid.ab[i] = 0.;
    if (has) {
      const double Te = (role == 0) ? 1.1 * S.T0 : ((role == 1) ? 0.9 * S.T0 : S.T0);
      balance_before_line_cooling(h0e, he0e, gain, mid, Te, ntot, midz, j, P.abund, h, P.tp.pahfac, S.crfac, P.tp.crscale,
                                  P.rr, out.x);
    }
    {
      /* line cooling: per lane while the warp is busy, warp-wide for its last pairs */
      const unsigned pairs = __ballot_sync(0xffffffffu, has);
      double cooling = 0.;
      if (__popc(pairs) > wide_max) {
        if (has) cooling = line_cooling(mid.T, mid.ne, mid.ab);
      } else {
        cooling = line_cooling_wide(s_lc_tab, pairs, has, mid.T, mid.ne, mid.ab);
      }
      if (has) balance_after_line_cooling(gain, loss, cooling, mid);
    }
    /* exchange within the slot (all lanes take part in the shuffles) */
    const double gain1 = __shfl_sync(0xffffffffu, gain, slot_base), loss1 = __shfl_sync(0xffffffffu, loss, slot_base);
    const double gain2 = __shfl_sync(0xffffffffu, gain, slot_base + 1), loss2 = __shfl_sync(0xffffffffu, loss, slot_base + 1);
    const int l2 = in_slot ? slot_base + 2 : lane;
    const double gain0 = __shfl_sync(0xffffffffu, gain, l2), loss0 = __shfl_sync(0xffffffffu, loss, l2);
    const double h00 = __shfl_sync(0xffffffffu, h0e, l2), he00 = __shfl_sync(0xffffffffu, he0e, l2);
    if (has) {
      ++S.niter;
      if (temperature_solve_update(S, gain1, loss1, gain2, loss2, h00, he00, gain0, loss0, P.tp)) {
        if (role == 2) { /* this lane holds the fractions of the evaluation at T0, the last one of the reference */
          /* temperature_solve_finish only looks at whether J_H / J_He are zero */
          double Jz[NUM_IONS];
#pragma unroll
          for (int k = 0; k < NUM_IONS; ++k) Jz[k] = 1.;
          if (jH_zero) Jz[ION_H_n] = 0.;
          if (jHe_zero) Jz[ION_He_n] = 0.;
          temperature_solve_finish(S, Jz, h, out);
          store(cell, ntot);
        }
        has = false;
      }
    }
  }
}

/* host SoA <-> device layout */
__global__ void pack_cells_kernel(int64_t ncell, const double *n, const double *T, const double *x,
                                  CellOpacity *cells, double2 *cells_h, double *xmetal) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  CellOpacity c;
  c.n = n[i];
  c.T = T[i];
  c.xH = x[i];
  c.xHe = x[ncell + i];
  cells[i] = c;
  cells_h[i] = make_double2(c.n, c.xH);
  for (int k = 0; k < 12; ++k) xmetal[i * 12 + k] = x[(2 + k) * ncell + i];
}

__global__ void unpack_cells_kernel(int64_t ncell, const CellOpacity *cells, const double *xmetal,
                                    const double *heat_norm, double *n, double *T, double *x,
                                    double *heat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const CellOpacity c = cells[i];
  n[i] = c.n;
  T[i] = c.T;
  x[i] = c.xH;
  x[ncell + i] = c.xHe;
  for (int k = 0; k < 12; ++k) x[(2 + k) * ncell + i] = xmetal[i * 12 + k];
  heat[i] = heat_norm[2 * i];
  heat[ncell + i] = heat_norm[2 * i + 1];
}

/* gathers of the owned chunks (UpdateParams): the owner packs its cells' records in work-item order, ncclAllGather
 * moves the equal-sized packs, every rank scatters the packs of the others back into cell order */
__global__ void pack_owned_records_kernel(int64_t n_work, int64_t ncells, int32_t rank, int32_t size, const CellOpacity *cells,
                                          CellOpacity *pack) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_work) return;
  const int64_t i = owned_cell(j, rank, size);
  CellOpacity c;
  c.n = c.xH = c.xHe = c.T = 0.;
  if (i < ncells) c = cells[i];
  pack[j] = c;
}
__global__ void unpack_owned_records_kernel(int64_t n_work, int64_t ncells, int32_t my_rank, int32_t size, const CellOpacity *packs,
                                            CellOpacity *cells, double2 *cells_h) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_work * size) return;
  const int32_t r = (int32_t)(t / n_work);
  if (r == my_rank) return;
  const int64_t i = owned_cell(t % n_work, r, size);
  if (i >= ncells) return;
  const CellOpacity c = packs[t];
  cells[i] = c;
  cells_h[i] = make_double2(c.n, c.xH);
}
/* the same for a plain array of `per_cell` doubles per cell (metal fractions, heating terms) */
__global__ void pack_owned_doubles_kernel(int64_t n_work, int64_t ncells, int32_t rank, int32_t size, int per_cell,
                                          const double *array, double *pack) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_work * per_cell) return;
  const int64_t i = owned_cell(t / per_cell, rank, size);
  pack[t] = (i < ncells) ? array[i * per_cell + t % per_cell] : 0.;
}
__global__ void unpack_owned_doubles_kernel(int64_t n_work, int64_t ncells, int32_t my_rank, int32_t size, int per_cell,
                                            const double *packs, double *array) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_work * per_cell * size) return;
  const int32_t r = (int32_t)(t / (n_work * per_cell));
  if (r == my_rank) return;
  const int64_t u = t % (n_work * per_cell);
  const int64_t i = owned_cell(u / per_cell, r, size);
  if (i < ncells) array[i * per_cell + u % per_cell] = packs[t];
}
/* host arrays that hold the cells a rank owns, in work-item order -> device layout (the distributed upload) */
__global__ void pack_cells_owned_kernel(int64_t n_owned, int32_t rank, int32_t size, const double *n, const double *T, const double *x,
                                        CellOpacity *cells, double2 *cells_h, double *xmetal) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_owned) return;
  const int64_t i = owned_cell(j, rank, size);
  CellOpacity c;
  c.n = n[j];
  c.T = T[j];
  c.xH = x[j];
  c.xHe = x[n_owned + j];
  cells[i] = c;
  cells_h[i] = make_double2(c.n, c.xH);
  for (int k = 0; k < 12; ++k) xmetal[i * 12 + k] = x[(2 + k) * n_owned + j];
}

/* host layout of the cells a rank owns, in work-item order (the distributed read-back of an iteration) */
__global__ void unpack_cells_owned_kernel(int64_t n_owned, int32_t rank, int32_t size, const CellOpacity *cells, const double *xmetal,
                                          const double *heat_norm, double *n, double *T, double *x, double *heat) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_owned) return;
  const int64_t i = owned_cell(j, rank, size);
  const CellOpacity c = cells[i];
  n[j] = c.n;
  T[j] = c.T;
  x[j] = c.xH;
  x[n_owned + j] = c.xHe;
  for (int k = 0; k < 12; ++k) x[(2 + k) * n_owned + j] = xmetal[i * 12 + k];
  heat[j] = heat_norm[2 * i];
  heat[n_owned + j] = heat_norm[2 * i + 1];
}

/* compact (n, x_H) copy of the opacity records of cells [lo, hi): rebuilt locally after a gather of the records */
__global__ void rebuild_cells_h_kernel(int64_t lo, int64_t hi, const CellOpacity *cells, double2 *cells_h) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= hi) return;
  const CellOpacity c = cells[i];
  cells_h[i] = make_double2(c.n, c.xH);
}

/* cmib_measure_scatter_rates: one FP64 RED / one 16-byte gather per lane and iteration, every lane in a different line */
CMIB_D uint32_t mb_hash32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
CMIB_D uint64_t mb_cell(uint32_t tid, uint32_t it, uint64_t ncell) {
  const uint32_t a = mb_hash32(tid * 0x9E3779B9u + it), b = mb_hash32(a ^ 0x85ebca6bu);
  return (((uint64_t)a << 32) | b) % ncell;
}
__global__ void measure_scatter_red_kernel(double *table, uint64_t ncell, int iters) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int it = 0; it < iters; ++it) atomicAdd(table + mb_cell(tid, it, ncell) * 2, 1.0);
}
__global__ void measure_scatter_gather_kernel(const double2 *table, uint64_t ncell, int iters, double *out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  double s = 0.;
  for (int it = 0; it < iters; ++it) {
    const double2 a = __ldg(table + mb_cell(tid, it, ncell));
    s += a.x + a.y;
  }
  if (s == 12345.678) out[0] = s;
}

/* accumulators -> reference SoA view J[14][ncell], heat[2][ncell] */
template <int MODE>
__global__ void unpack_acc_kernel(int64_t ncell, const double *acc, int64_t honly_cell_stride,
                                  int64_t honly_term_stride, int64_t honly_offset, double *J, double *heat) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ncell) return;
  const double *a = acc + ACC_COUNTERS + i * AccLayout<MODE>::NACC;
  if (MODE == ACC_HONLY) {
    for (int k = 1; k < NUM_IONS; ++k) J[k * ncell + i] = 0.;
    J[i] = acc[ACC_COUNTERS + honly_offset + i * honly_cell_stride];
    heat[i] = acc[ACC_COUNTERS + honly_offset + i * honly_cell_stride + honly_term_stride];
    heat[ncell + i] = 0.;
  } else {
    for (int k = 0; k < NUM_IONS; ++k) J[k * ncell + i] = a[acc_slot(k)];
    heat[i] = a[acc_slot(NUM_IONS)];
    heat[ncell + i] = a[acc_slot(NUM_IONS + 1)];
  }
}

/* ------------------------------------------------------------------------- */
/* element-wise probes                                                        */
/* ------------------------------------------------------------------------- */
__global__ void eval_cross_sections_kernel(int64_t n, SourceModel m, const double *nu, double *sigma) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double s[NUM_IONS], sHe;
  packet_cross_sections<NUM_IONS>(m, nu[i], s, sHe);
  for (int k = 0; k < NUM_IONS; ++k) sigma[i * NUM_IONS + k] = s[k];
}

__global__ void eval_recombination_kernel(int64_t n, RecombinationModel rr, const double *T,
                                          double *alpha) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  for (int k = 0; k < NUM_IONS; ++k) alpha[i * NUM_IONS + k] = recombination_rate(rr, k, T[i]);
}

__global__ void eval_charge_transfer_kernel(int64_t n, const double *T4, double *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double *o = out + i * 3 * NUM_IONS;
  for (int k = 0; k < NUM_IONS; ++k) {
    o[k] = (k == ION_H_n) ? 0. : ct_recombination_H(k, T4[i]);
    o[NUM_IONS + k] = ct_ionization_H(k, T4[i]);
    o[2 * NUM_IONS + k] = ct_recombination_He(k, T4[i]);
  }
}

__global__ void eval_line_cooling_kernel(int64_t n, const double *T, const double *ne,
                                         const double *abund, double *cooling) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double ab[LC_NUM];
  for (int k = 0; k < LC_NUM; ++k) ab[k] = abund[i * LC_NUM + k];
  cooling[i] = line_cooling(T[i], ne[i], ab);
}

__global__ void eval_solve5_kernel(int64_t n, double *A, double *B, int32_t *status) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double a[5][5], b[5];
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int c = 0; c < 5; ++c) a[r][c] = A[i * 25 + r * 5 + c];
    b[r] = B[i * 5 + r];
  }
  status[i] = solve5(a, b);
#pragma unroll
  for (int r = 0; r < 5; ++r) {
#pragma unroll
    for (int c = 0; c < 5; ++c) A[i * 25 + r * 5 + c] = a[r][c];
    B[i * 5 + r] = b[r];
  }
}

__global__ void eval_reemission_probabilities_kernel(int64_t n, const double *T, double *out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double p[NUM_REEMIT];
  reemission_probabilities(T[i], p);
  for (int k = 0; k < NUM_REEMIT; ++k) out[i * NUM_REEMIT + k] = p[k];
}

struct EvalStateParams {
  int64_t n;
  double jfac, hfac;
  double abund[NUM_ELEMENTS];
  RecombinationModel rr;
  TemperatureParams tp;
  const double *J, *heat, *ndens, *T, *cr_factor, *midz;
  double *T_out, *x, *heat_out;
  int solve_temperature;
};

__global__ void __launch_bounds__(128) eval_state_kernel(const __grid_constant__ EvalStateParams P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  double J[NUM_IONS], heat[NUM_HEAT];
  for (int k = 0; k < NUM_IONS; ++k) J[k] = P.J[k * P.n + i];
  heat[0] = P.heat[i];
  heat[1] = P.heat[P.n + i];
  CellState out;
  if (P.solve_temperature) {
    double xprev[NUM_IONS];
    for (int k = 0; k < NUM_IONS; ++k) xprev[k] = 0.;
    cell_temperature(P.jfac, P.hfac, J, heat, P.ndens[i], P.T[i], P.cr_factor ? P.cr_factor[i] : -1.,
                     P.midz ? P.midz[i] : 0., P.abund, P.rr, P.tp, xprev, out);
  } else {
    cell_ionization_state(P.jfac, P.hfac, J, heat, P.ndens[i], P.T[i], P.abund, P.rr, out);
  }
  if (P.T_out) P.T_out[i] = out.T;
  for (int k = 0; k < NUM_IONS; ++k) P.x[k * P.n + i] = out.x[k];
  P.heat_out[i] = out.heat[0];
  P.heat_out[P.n + i] = out.heat[1];
}

struct EvalBalanceParams {
  int64_t n;
  double abund[NUM_ELEMENTS];
  RecombinationModel rr;
  double pahfac, crfac, crscale;
  const double *T, *ndens, *j, *h, *midz;
  double *h0, *he0, *gain, *loss, *metals;
};

__global__ void __launch_bounds__(128) eval_balance_kernel(const __grid_constant__ EvalBalanceParams P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  double j[NUM_IONS], h[NUM_HEAT], x[NUM_IONS];
  for (int k = 0; k < NUM_IONS; ++k) { j[k] = P.j[i * NUM_IONS + k]; x[k] = 0.; }
  h[0] = P.h[2 * i];
  h[1] = P.h[2 * i + 1];
  double h0, he0, gain, loss;
  cooling_heating_balance(h0, he0, gain, loss, P.T[i], P.ndens[i], P.midz ? P.midz[i] : 0., j,
                          P.abund, h, P.pahfac, P.crfac, P.crscale, P.rr, x);
  P.h0[i] = h0; P.he0[i] = he0; P.gain[i] = gain; P.loss[i] = loss;
  for (int k = 0; k < 12; ++k) P.metals[i * 12 + k] = x[2 + k];
}

struct SamplePacketsParams {
  SourceModel src;
  GridGeom geom;
  int64_t n;
  uint64_t offset, seed;
  uint32_t iteration;
  double *pos, *dir, *nu, *sigma, *sigma_He_corr, *tau;
};

__global__ void sample_packets_kernel(const __grid_constant__ SamplePacketsParams P) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= P.n) return;
  const SourceModel &m = P.src;
  PacketRng rng;
  rng_init(rng, P.seed, P.iteration, P.offset + i);
  double px, py, pz, dx, dy, dz, nu;
  int isrc;
  emit_primary(m, P.geom, rng, px, py, pz, dx, dy, dz, nu, isrc);
  double s[NUM_IONS], sHe;
  packet_cross_sections<NUM_IONS>(m, nu, s, sHe);
  const double tau = -log(rng_uniform(rng));
  P.pos[3 * i] = px; P.pos[3 * i + 1] = py; P.pos[3 * i + 2] = pz;
  P.dir[3 * i] = dx; P.dir[3 * i + 1] = dy; P.dir[3 * i + 2] = dz;
  P.nu[i] = nu;
  for (int k = 0; k < NUM_IONS; ++k) P.sigma[i * NUM_IONS + k] = s[k];
  P.sigma_He_corr[i] = sHe;
  P.tau[i] = tau;
}

__global__ void sample_spectrum_kernel(SourceModel m, int which, double T, uint64_t seed, int64_t n,
                                       double *nu) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  PacketRng rng;
  rng_init(rng, seed, 0u, (uint64_t)i);
  double v;
  if (which == 0) v = spectrum_frequency(m.spectrum, rng);
  else if (which == 4) v = spectrum_frequency(m.cont_spectrum, rng);
  else if (which == 1) v = lyc_frequency(m.hlyc_freq, m.hlyc_temp, m.hlyc_cdf, T, rng, m.hlyc_guide);
  else if (which == 2) v = lyc_frequency(m.helyc_freq, m.helyc_temp, m.helyc_cdf, T, rng, m.helyc_guide);
  else v = he2pc_frequency(m.he2pc_freq, m.he2pc_cdf, rng, m.he2pc_guide);
  nu[i] = v;
}

} // namespace cmib
