/*
 * cmib_api.cu — implementation of the C ABI declared in include/cmib.h.
 *
 * Host side is thin: it owns device buffers, builds the tabulated spectra on the
 * host (spectrum_tables.hpp), and launches the kernels of kernels.cuh on the
 * context's stream.  There is no CPU compute path: every entry point that does
 * work requires a CUDA device and fails loudly otherwise.
 */
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/cmib.h"
#include "kernels.cuh"
#include "wavefront.cuh"
#include "march_coherent.cuh"
#include "spectrum_tables.hpp"

using namespace cmib;

namespace {

thread_local std::string g_last_error;
int g_abort_on_error = -1; /* -1: read CMIB_ABORT_ON_ERROR lazily */
std::atomic<uint64_t> g_launches{0}; /* contexts may be driven from several host threads */

int fail(const char *file, const char *func, int line, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  char full[1200];
  /* same shape as the reference's cmac_error (src/Error.hpp:101-106) */
  snprintf(full, sizeof(full), "%s:%s():%d: Error: %s", file, func, line, buf);
  g_last_error = full;
  if (g_abort_on_error < 0) {
    const char *e = getenv("CMIB_ABORT_ON_ERROR");
    g_abort_on_error = (e && e[0] == '1') ? 1 : 0;
  }
  if (g_abort_on_error) {
    fprintf(stderr, "%s\n", full);
    abort();
  }
  return 1;
}
#define CMIB_FAIL(...) return fail(__FILE__, __func__, __LINE__, __VA_ARGS__)
#define CUDA_OK(expr)                                                                   \
  do {                                                                                  \
    cudaError_t e_ = (expr);                                                            \
    if (e_ != cudaSuccess) CMIB_FAIL("%s failed: %s", #expr, cudaGetErrorString(e_));   \
  } while (0)
#define CHECK_CTX(ctx)                                                                  \
  do {                                                                                  \
    if (!(ctx)) CMIB_FAIL("null context");                                              \
    CUDA_OK(cudaSetDevice((ctx)->device));                                              \
  } while (0)

template <typename T> struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  cudaError_t resize(size_t count) {
    if (count == n) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (count == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc(&p, count * sizeof(T));
    if (e == cudaSuccess) n = count;
    return e;
  }
  cudaError_t upload(const T *h, size_t count, cudaStream_t s) {
    cudaError_t e = resize(count);
    if (e != cudaSuccess) return e;
    return cudaMemcpyAsync(p, h, count * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  ~DevBuf() {
    if (p) cudaFree(p);
  }
};

inline unsigned blocks_for(int64_t n, int bs) { return (unsigned)((n + bs - 1) / bs); }

} // namespace

struct cmib_context {
  int device = 0;
  int sm_count = 0;
  cudaStream_t stream = nullptr;
  GridGeom geom;
  /* grid state */
  DevBuf<CellOpacity> cells;
  DevBuf<double2> cells_h;
  DevBuf<double> xmetal, heat_norm, cr_factor, reemit_prob, acc;
  bool have_cr_factor = false;
  bool reemit_prob_valid = false;
  /* staging for SoA <-> device layout conversion */
  DevBuf<double> stage;
  /* plugins */
  double abund[NUM_ELEMENTS] = {0., 0., 0., 0., 0., 0.};
  SourceModel src;
  RecombinationModel rr;
  TemperatureParams tp;
  double luminosity = 0.; /* discrete + continuous */
  double discrete_luminosity = 0., continuous_luminosity = 0.;
  bool planar_geometry_set = false, star_position_set = false, disc_geometry_set = false, galaxy_geometry_set = false;
  DevBuf<double> d_galaxy_tables; /* [2][GALAXY_NBIN + 1]: radius, cumulative disc luminosity */
  DevBuf<double> d_cont_planck;
  DevBuf<uint16_t> d_cont_planck_guide;
  std::vector<double> h_cont_planck;
  DevBuf<double> d_spec_freq[2], d_spec_cdf[2]; /* tabulated spectra: [0] discrete sources, [1] continuous source */
  /* PhotonSource.cpp:110-131: probability of a continuous packet and the two packet weights */
  void update_source_weights() {
    luminosity = discrete_luminosity + continuous_luminosity;
    if (src.continuous_kind != CONTINUOUS_NONE && continuous_luminosity > 0.) {
      if (src.n_sources > 0 && discrete_luminosity > 0.) {
        src.continuous_probability = 0.5;
        src.discrete_weight = 1.;
        src.continuous_weight = (1. - src.continuous_probability) * continuous_luminosity /
                                src.continuous_probability / discrete_luminosity;
      } else {
        src.continuous_probability = 1.;
        src.discrete_weight = 0.;
        src.continuous_weight = 1.;
      }
    } else {
      src.continuous_probability = 0.;
      src.discrete_weight = 1.;
      src.continuous_weight = 0.;
    }
  }
  DevBuf<uint16_t> d_planck_guide, d_hlyc_guide, d_helyc_guide, d_he2pc_guide;
  DevBuf<double> d_src_pos, d_src_cum, d_planck, d_hlyc_freq, d_hlyc_temp, d_hlyc_cdf, d_helyc_freq,
      d_helyc_temp, d_helyc_cdf, d_he2pc_freq, d_he2pc_cdf;
  std::vector<double> h_planck, h_hlyc_freq, h_hlyc_temp, h_hlyc_cdf, h_helyc_freq, h_helyc_temp,
      h_helyc_cdf, h_he2pc_freq, h_he2pc_cdf;
  int acc_mode = ACC_FULL;
  bool force_full = false;
  /* wavefront shoot (wavefront.cuh): queues + control block, allocated on first use */
  int shoot_algorithm = 0; /* 0 wavefront (production), 1 one-thread-per-packet kernel (A/B check) */
  uint64_t queue_capacity = 0;
  int queue_mode = -1;
  DevBuf<double> mq, rq, eq;
  DevBuf<unsigned long long> ctl;
  DevBuf<uint32_t> sort_key, sort_order, sort_hist;
  DevBuf<uint32_t> sort_key_out, sort_iota; /* coherent march (sort mode 2): radix sort of (key, slot) pairs */
  DevBuf<unsigned char> sort_temp;
  DevBuf<double> hot_acc;       /* replicated accumulators of the cells around the sources */
  DevBuf<uint32_t> d_src_cell;  /* packed cell indices of the sources */
  std::vector<uint32_t> h_src_cell;
  int hot_replicas = 0;
  DevBuf<unsigned long long> upd_counter; /* next unprocessed cell of update_temperature_kernel */
  int update_blocks_per_sm[2] = {0, 0};
  /* march-queue order: 0 emission order, 1 coarse counting sort (measured slower),
   * 2 coherent march = fine radix sort + in-warp sums (march_kernel<MODE, true>),
   * -1 (default) measured: grids that fit in L2 use 0 (2 loses there on every workload measured);
   * for larger grids the two are timed on successive large shoots of this context — time per cell
   * crossing, so that the growth of the ionised volume from one shoot to the next does not bias the
   * comparison — and the faster one is kept (a shoot of >= 4 queue capacities times the two orders on
   * its own rounds 1 and 2 instead and finishes in the winner).  The timed pair is repeated after 2, 4, 8 and
   * then every 16 shoots: the regime changes quickly during the first iterations of a run (a small
   * ionised bubble fits in L2, the converged one may not).  Both orders shoot the same packets; only
   * the order of the atomic adds differs. */
  int sort_mode = -1;
  int tune_shoots = 0;                 /* large shoots seen */
  int tune_next = 1;                   /* shoot at which the next timed pair starts */
  int tune_interval = 2;               /* shoots between the end of a pair and the next one */
  double tune_ns_per_crossing[3] = {0., 0., 0.}; /* indexed by order (0, 2) */
  int tuned_order = 0;
  cudaEvent_t tune_ev[3] = {nullptr, nullptr, nullptr};
  size_t l2_bytes = 0;
  unsigned long long *h_ctl = nullptr; /* pinned mirror of the control block */
  int march_blocks_per_sm[2][2] = {{0, 0}, {0, 0}}; /* [layout][plain, coherent] */
  int lean_grid[2][2][2][2] = {};                   /* resident CTAs per SM of the march_lean_kernel variants */
  int prep_blocks_per_sm[2] = {0, 0};
  int decide_blocks_per_sm = 0;
  uint64_t shoot_rounds = 0;
  /* optional per-kernel timing of the shoot (CUDA events on the context's stream) */
  bool timing = false;
  std::vector<cudaEvent_t> ev_pool;
  double prepare_ms = 0., march_ms = 0.;
  double nu_H = 0., nu_He = 0.;

  int pick_acc_mode() const {
    if (force_full) return ACC_FULL;
    if (src.xs_kind == XS_VERNER) return ACC_FULL;
    for (int k = 1; k < NUM_IONS; ++k)
      if (src.xs_fixed[k] != 0. || (src.xs_kind == XS_BIMODAL && src.xs_high[k] != 0.)) return ACC_FULL;
    return ACC_HONLY;
  }
  /* H-only accumulator layout: planes when cells_h + accumulators do not fit in L2 */
  bool honly_planar() const { return (size_t)geom.ncells * 32 > l2_bytes; }
  /* L2-resident grids: one cell per 128-B line — atomics to different cells of one line
   * serialise in L2 (measured on stromgren 64^3: 2.24 ms at 16 B/cell, 1.95 ms at 32 B/cell,
   * 1.67 ms at 128 B/cell per 4e6 packets); in between: interleaved J, heat */
  int64_t honly_cell_stride() const {
    if (honly_planar()) return 1;
    if (const char *e = getenv("CMIB_HONLY_STRIDE")) return atoi(e) >= 2 ? atoi(e) : 2;
    return ((size_t)geom.ncells * 128 <= l2_bytes / 2) ? 16 : 2;
  }
  int64_t honly_term_stride() const { return honly_planar() ? geom.ncells : 1; }
  /* padded records: J_H, heat_H sit in the second half of the cell's 128-B line.  Measured on
   * stromgren 64^3 (4e6 packets): march 1.49 ms with the pair at byte 64 of the line, 1.78 ms at
   * byte 0 (profiles/r01_layout_experiments.md) */
  int64_t honly_offset() const {
    if (const char *e = getenv("CMIB_HONLY_OFFSET")) return atoi(e);
    return (!honly_planar() && honly_cell_stride() == 16) ? 8 : 0;
  }
  size_t acc_doubles(int mode) const {
    if (mode != ACC_HONLY) return ACC_COUNTERS + (size_t)geom.ncells * 16;
    return ACC_COUNTERS + 64 + (size_t)geom.ncells * (honly_planar() ? 2 : (size_t)honly_cell_stride());
  }
};

namespace {

/* the accumulator layout follows the cross-section model; switching layouts
 * discards the current sums (they are per-iteration scratch anyway) */
int ensure_acc(cmib_context *ctx) {
  const int mode = ctx->pick_acc_mode();
  const size_t want = ctx->acc_doubles(mode);
  if (mode != ctx->acc_mode || ctx->acc.n != want) {
    ctx->acc_mode = mode;
    CUDA_OK(ctx->acc.resize(want));
    CUDA_OK(cudaMemsetAsync(ctx->acc.p, 0, want * sizeof(double), ctx->stream));
  }
  return 0;
}

/* build the bracket guides of `rows` CDF rows, check on the host that the guided search returns
 * exactly what Utilities::locate returns, upload */
int make_guides(cmib_context *ctx, const double *cdf, int rows, DevBuf<uint16_t> &dev, const uint16_t **out) {
  std::vector<uint16_t> g((size_t)rows * (GUIDE_N + 1));
  for (int r = 0; r < rows; ++r)
    host::build_guide(cdf + (size_t)r * SPECTRUM_NUMFREQ, SPECTRUM_NUMFREQ, GUIDE_N, g.data() + (size_t)r * (GUIDE_N + 1));
  for (int r = 0; r < rows; ++r) {
    const double *row = cdf + (size_t)r * SPECTRUM_NUMFREQ;
    const uint16_t *gr = g.data() + (size_t)r * (GUIDE_N + 1);
    for (int k = 0; k <= 4096; ++k) {
      const double x = (k + 0.37) / 4097.;
      if (locate_guided(x, row, SPECTRUM_NUMFREQ, gr) != locate(x, row, SPECTRUM_NUMFREQ))
        CMIB_FAIL("internal error: guided CDF search disagrees with bisection (row %d, x %g)", r, x);
    }
    for (int j = 0; j < SPECTRUM_NUMFREQ; ++j) { /* exactly on the table values, and one ulp around them */
      const double xs[3] = {row[j], nextafter(row[j], 0.), nextafter(row[j], 2.)};
      for (double x : xs)
        if (x >= 0. && x <= 1. && locate_guided(x, row, SPECTRUM_NUMFREQ, gr) != locate(x, row, SPECTRUM_NUMFREQ))
          CMIB_FAIL("internal error: guided CDF search disagrees with bisection (row %d, entry %d)", r, j);
    }
  }
  CUDA_OK(dev.upload(g.data(), g.size(), ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  *out = dev.p;
  return 0;
}

/* UnitConverter::to_SI<QUANTITY_FREQUENCY>(x, "eV") (UnitConverter.hpp:259-275):
 * value * eV, then times (1/h) */
double ev_to_hz(double ev) { return (ev * ELECTRONVOLT) * (1. / PLANCK); }

} // namespace

/* scratch device buffers of the test hooks */
namespace {
struct Scratch {
  std::vector<void *> ptrs;
  ~Scratch() {
    for (void *p : ptrs) cudaFree(p);
  }
  template <typename T> cudaError_t in(T **d, const T *h, size_t n, cudaStream_t s) {
    *d = nullptr;
    if (!h || n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void **)d, n * sizeof(T));
    if (e != cudaSuccess) return e;
    ptrs.push_back(*d);
    return cudaMemcpyAsync(*d, h, n * sizeof(T), cudaMemcpyHostToDevice, s);
  }
  template <typename T> cudaError_t out(T **d, size_t n) {
    *d = nullptr;
    if (n == 0) return cudaSuccess;
    cudaError_t e = cudaMalloc((void **)d, n * sizeof(T));
    if (e == cudaSuccess) ptrs.push_back(*d);
    return e;
  }
  template <typename T> cudaError_t back(T *h, const T *d, size_t n, cudaStream_t s) {
    if (!h || !d || n == 0) return cudaSuccess;
    return cudaMemcpyAsync(h, d, n * sizeof(T), cudaMemcpyDeviceToHost, s);
  }
};
} // namespace


namespace {

/* PhotonSourceSpectrumFactory (src/PhotonSourceSpectrumFactory.hpp:84-152) for the closed-form and
 * Planck spectra; tabulated ones come through cmib_set_spectrum_table */
int set_spectrum_model(cmib_context *ctx, SpectrumModel &sp, std::vector<double> &h_planck, DevBuf<double> &d_planck,
                       DevBuf<uint16_t> &d_guide, int kind, double param) {
  if (kind == CMIB_SPECTRUM_MONOCHROMATIC) {
    sp.kind = SPECTRUM_MONOCHROMATIC;
    sp.mono_frequency = param;
  } else if (kind == CMIB_SPECTRUM_PLANCK) {
    if (!(param > 0.)) CMIB_FAIL("Planck temperature must be positive");
    host::build_planck_table(param, h_planck);
    CUDA_OK(d_planck.upload(h_planck.data(), h_planck.size(), ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    sp.kind = SPECTRUM_PLANCK;
    sp.planck = d_planck.p;
    if (make_guides(ctx, h_planck.data(), 1, d_guide, &sp.planck_guide)) return 1;
  } else if (kind == CMIB_SPECTRUM_UNIFORM) {
    sp.kind = SPECTRUM_UNIFORM;
  } else {
    CMIB_FAIL("Unknown PhotonSourceSpectrum type: %d", kind);
  }
  return 0;
}

uint64_t default_queue_capacity() {
  const char *e = getenv("CMIB_QUEUE_CAPACITY");
  if (e) {
    const long long v = atoll(e);
    if (v >= 1024) return (uint64_t)v;
  }
  return 1ull << 24; /* 16 Mi packets: 5 GB of queues (full layout); measured 124 -> 117 ms per lexingtonHII20 step vs 4 Mi */
}

/* one cmib_shoot call on the wavefront path: rounds of prepare -> march until the
 * queues run dry.  The host only reads back the march-queue sizes every few rounds. */
int shoot_wavefront(cmib_context *ctx, const ShootParams &P) {
  const int mode = ctx->acc_mode;
  uint64_t cap = default_queue_capacity();
  if (P.n_packets < cap) cap = (P.n_packets + 1023) / 1024 * 1024;
  const int nf = (mode == ACC_HONLY) ? MarchQueueLayout<ACC_HONLY>::NFIELDS : MarchQueueLayout<ACC_FULL>::NFIELDS;
  if (ctx->queue_capacity < cap || ctx->queue_mode != mode) {
    CUDA_OK(ctx->mq.resize((size_t)nf * cap));
    CUDA_OK(ctx->rq.resize((size_t)RQ_NFIELDS * cap));
    CUDA_OK(ctx->eq.resize((size_t)EQ_NFIELDS * cap));
    ctx->queue_capacity = cap;
    ctx->queue_mode = mode;
  }
  cap = ctx->queue_capacity;
  if (!ctx->ctl.p) {
    CUDA_OK(ctx->ctl.resize(CTL_WORDS));
    CUDA_OK(cudaMallocHost((void **)&ctx->h_ctl, CTL_WORDS * sizeof(unsigned long long)));
    int occ[4] = {0, 0, 0, 0};
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[0], march_kernel<ACC_FULL, false, true>, MARCH_BLOCK, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[1], march_kernel<ACC_FULL, true, true>, MARCH_BLOCK, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[2], march_kernel<ACC_HONLY, false, true>, MARCH_BLOCK, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ[3], march_kernel<ACC_HONLY, true, true>, MARCH_BLOCK, 0));
    ctx->march_blocks_per_sm[ACC_FULL][0] = occ[0];
    ctx->march_blocks_per_sm[ACC_FULL][1] = occ[1];
    ctx->march_blocks_per_sm[ACC_HONLY][0] = occ[2];
    ctx->march_blocks_per_sm[ACC_HONLY][1] = occ[3];
  }
  cudaStream_t s = ctx->stream;
  memset(ctx->h_ctl, 0, CTL_WORDS * sizeof(unsigned long long));
  ctx->h_ctl[CTL_REMAINING] = P.n_packets;
  CUDA_OK(cudaMemcpyAsync(ctx->ctl.p, ctx->h_ctl, CTL_WORDS * sizeof(unsigned long long), cudaMemcpyHostToDevice, s));
  WavefrontParams W;
  W.sp = P;
  W.ctl = ctx->ctl.p;
  W.mq = ctx->mq.p;
  W.rq = ctx->rq.p;
  W.eq = ctx->eq.p;
  W.capacity = cap;
  W.acc_j = P.acc + ACC_COUNTERS + P.honly_offset;
  for (int d = 0; d < 3; ++d) W.lean_n16[d] = 16u * (uint32_t)P.geom.ncell[d];
  W.lean_k[0] = P.geom.ncell[1] * P.geom.ncell[2];
  W.lean_k[1] = P.geom.ncell[2];
  /* order of the march queue (cmib_context::sort_mode; CMIB_SORT overrides for A/B runs) */
  int sort = ctx->sort_mode;
  if (const char *e = getenv("CMIB_SORT")) sort = atoi(e);
  bool tuning = false;        /* this whole shoot is one half of a timed pair */
  bool tune_in_shoot = false; /* rounds 1 and 2 of this shoot are the timed pair */
  if (sort < 0) {
    const size_t working_set = (size_t)ctx->geom.ncells * ((mode == ACC_HONLY ? 16 : sizeof(CellOpacity)) + (mode == ACC_HONLY ? 16 : 128));
    if (working_set <= ctx->l2_bytes || P.n_packets < (1ull << 20)) {
      sort = (working_set <= ctx->l2_bytes) ? 0 : ctx->tuned_order;
    } else {
      /* shoot 0 is a warm-up (allocations); a timed pair = order 0, then order 2 */
      const int phase = ctx->tune_shoots - ctx->tune_next;
      /* a shoot of at least four full queues carries the pair itself: its rounds 1 and 2 (same grid
       * state, same mix of primaries and re-emitted packets) run in order 0 and in order 2, the rest
       * in the winner — one round of ~n/capacity in the slower order instead of a whole shoot */
      tune_in_shoot = (phase == 0 && P.n_packets >= 4 * cap);
      sort = tune_in_shoot ? 2 : ((phase == 0) ? 0 : (phase == 1 ? 2 : ctx->tuned_order));
      tuning = !tune_in_shoot && (phase == 0 || phase == 1);
      if (phase == 1 || tune_in_shoot) {
        ctx->tune_next = ctx->tune_shoots + 1 + ctx->tune_interval;
        if (ctx->tune_interval < 16) ctx->tune_interval *= 2;
      }
      ++ctx->tune_shoots;
    }
  }
  W.sort = sort;
  W.key = nullptr; W.order = nullptr; W.hist = nullptr; W.nbins = 0; W.isrc_bits_shift = SORT_DIR_BITS;
  W.sort_n = 0; W.fine_dir_bits = 22; W.fine_key_bits = 22; W.chunk_stride = 1; W.agg = 1;
  size_t sort_temp_bytes = 0;
  bool sort_reemitted_rounds = false;
  if (sort == 2) {
    /* key = source | direction (wavefront.cuh): as many direction bits as the source index leaves */
    int src_bits = 0;
    while ((1ll << src_bits) < (long long)P.src.n_sources) ++src_bits;
    /* 24 key bits = 3 radix passes while that leaves >= 18 direction bits (512 x 512 bins per source) */
    int dir_bits = 23 - src_bits;
    if (dir_bits > 22) dir_bits = 22;
    if (dir_bits < 18) dir_bits = 18;
    if (dir_bits + src_bits > 30) dir_bits = 30 - src_bits;
    if (dir_bits < 2) W.sort = sort = 0; /* more than 2^28 sources: no room for a direction */
    W.fine_dir_bits = dir_bits & ~1;
    W.fine_key_bits = W.fine_dir_bits + src_bits;
  }
  if (sort == 2) {
    W.sort = 2;
    if (const char *e = getenv("CMIB_CHUNK_STRIDE")) W.chunk_stride = (uint32_t)atoll(e); /* e.g. the prime 1000003 */
    if (W.chunk_stride < 1u || cap / MARCH_CHUNK >= W.chunk_stride) W.chunk_stride = 1u;
    if (const char *e = getenv("CMIB_AGG")) W.agg = atoi(e) != 0;
    if (const char *e = getenv("CMIB_SORT_REEMITTED")) sort_reemitted_rounds = atoi(e) != 0;
    if (ctx->sort_key.n < cap) {
      CUDA_OK(ctx->sort_key.resize(cap));
      CUDA_OK(ctx->sort_order.resize(cap));
    }
    if (ctx->sort_key_out.n < cap) {
      CUDA_OK(ctx->sort_key_out.resize(cap));
      CUDA_OK(ctx->sort_iota.resize(cap));
      iota_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->sort_iota.p, cap);
    }
    CUDA_OK(cub::DeviceRadixSort::SortPairs(nullptr, sort_temp_bytes, ctx->sort_key.p, ctx->sort_key_out.p,
                                            ctx->sort_iota.p, ctx->sort_order.p, (int64_t)cap, 0, W.fine_key_bits + 1, ctx->stream));
    if (ctx->sort_temp.n < sort_temp_bytes) CUDA_OK(ctx->sort_temp.resize(sort_temp_bytes));
    sort_temp_bytes = ctx->sort_temp.n;
    W.key = ctx->sort_key.p; W.order = ctx->sort_order.p;
  } else if (sort) {
    /* keep the number of bins <= 2^20: fewer direction bits when there are many sources */
    int shift = 6; /* 8 x 8 direction bins per source: see wavefront.cuh */
    if (const char *e = getenv("CMIB_SORT_DIR_BITS")) shift = atoi(e);
    if (shift < 0) shift = 0;
    if (shift > SORT_DIR_BITS) shift = SORT_DIR_BITS;
    while (shift > 0 && ((uint64_t)P.src.n_sources << shift) > (1ull << 20)) --shift;
    W.isrc_bits_shift = shift;
    W.nbins = (uint32_t)(((uint64_t)P.src.n_sources << shift) + SORT_POS_BINS);
    if (ctx->sort_key.n < cap) {
      CUDA_OK(ctx->sort_key.resize(cap));
      CUDA_OK(ctx->sort_order.resize(cap));
    }
    if (ctx->sort_hist.n < W.nbins + 1) CUDA_OK(ctx->sort_hist.resize(W.nbins + 1));
    W.key = ctx->sort_key.p; W.order = ctx->sort_order.p; W.hist = ctx->sort_hist.p;
  }
  /* persistent grids: exactly the CTAs that are resident at once (a partial second wave of a
   * grid-stride kernel runs at a fraction of the machine) */
  if (ctx->prep_blocks_per_sm[0] == 0) {
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prep_blocks_per_sm[ACC_FULL], prepare_kernel<ACC_FULL>, 256, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prep_blocks_per_sm[ACC_HONLY], prepare_kernel<ACC_HONLY>, 256, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->decide_blocks_per_sm, reemit_decide_kernel, 256, 0));
  }
  const unsigned prep_grid = (unsigned)(ctx->sm_count * (ctx->prep_blocks_per_sm[mode] > 0 ? ctx->prep_blocks_per_sm[mode] : 1));
  const unsigned decide_grid = (unsigned)(ctx->sm_count * (ctx->decide_blocks_per_sm > 0 ? ctx->decide_blocks_per_sm : 1));
  unsigned march_grids[2];
  for (int a = 0; a < 2; ++a) {
    int bpm = ctx->march_blocks_per_sm[mode][a];
    if (const char *e = getenv("CMIB_MARCH_BLOCKS_PER_SM")) bpm = atoi(e);
    if (bpm < 1) bpm = 1;
    march_grids[a] = (unsigned)(ctx->sm_count * bpm);
  }
  const int group = 4;
  uint64_t round = 0;
  size_t ev_used = 0;
  auto stamp = [&]() {
    if (!ctx->timing) return;
    if (ev_used == ctx->ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ctx->ev_pool.push_back(e);
    }
    cudaEventRecord(ctx->ev_pool[ev_used++], s);
  };
  ctx->prepare_ms = ctx->march_ms = 0.;
  /* upper bound: every round either emits min(remaining, room) primaries or shrinks the
   * re-emission population; 1e6 rounds cannot be reached by a sane configuration */
  /* coherent march: what the host knows about the coming rounds (read back once per group):
   * while primaries remain a round can fill the queue; afterwards it holds at most the packets
   * of the round before.  Rounds without primaries (re-emitted packets start anywhere) run
   * unsorted through the plain kernel. */
  const int sort_cfg = W.sort;
  /* H-only coherent walk: march_lean_kernel (CMIB_LEAN=0: the r01 kernel, for A/B runs) */
  int lean_cfg = 1, lean_steps = 3;
  if (const char *e = getenv("CMIB_LEAN")) lean_cfg = atoi(e) != 0;
  if (const char *e = getenv("CMIB_LEAN_STEPS")) lean_steps = atoi(e) == 2 ? 2 : 3;
  const size_t lean_smem = lean_smem_bytes(P.geom);
  if (lean_smem > 200 * 1024) lean_cfg = 0; /* wall tables of > ~11000 cells per axis sum: the r01 kernel */
  /* a heat term exists unless every packet of the shoot sits exactly at the threshold: a monochromatic
   * source at nu_H without re-emission (then nu - nu_H == 0 and the r01 kernel skipped the add at run time) */
  const int lean_heat = !(P.src.spectrum.kind == SPECTRUM_MONOCHROMATIC && P.src.spectrum.mono_frequency == P.nu_H &&
                          P.src.reemission_kind == REEMISSION_NONE && P.src.continuous_kind == CONTINUOUS_NONE);
  const int lean_periodic = (P.geom.periodic[0] | P.geom.periodic[1] | P.geom.periodic[2]) ? 1 : 0;
  /* march_kernel<.., PRE>: request the next cell record one pass ahead.  Measured (B200): pays in the
   * coherent kernel, whose in-warp sums sit between request and use (clumpy 256^3 30.1 -> 26.3 ms);
   * the plain kernel is bound by L1TEX lanes, not latency (no gain; -16 % with the full layout's spills) */
  int prefetch_cfg = -1;
  if (const char *e = getenv("CMIB_PREFETCH")) prefetch_cfg = atoi(e) != 0;
  double tune_crossings0 = 0.;
  if (tuning) { /* buffers are allocated; wait for earlier work: time the shoot alone */
    CUDA_OK(cudaMemcpyAsync(&tune_crossings0, ctx->acc.p + 5, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
  }
  const auto tune_t0 = std::chrono::steady_clock::now();
  uint64_t items_bound = cap;
  bool primaries_left = true;
  int order_now = tune_in_shoot ? ctx->tuned_order : sort_cfg;
  if (tune_in_shoot && !ctx->tune_ev[0])
    for (int k = 0; k < 3; ++k) CUDA_OK(cudaEventCreate(&ctx->tune_ev[k]));
  while (round < 1000000) {
    for (int k = 0; k < group; ++k, ++round) {
      if (tune_in_shoot && round >= 1 && round <= 3) CUDA_OK(cudaEventRecord(ctx->tune_ev[round - 1], s));
      if (sort_cfg == 2) {
        W.sort = (order_now == 2 && (primaries_left || sort_reemitted_rounds)) ? 2 : 0;
        if (tune_in_shoot && round == 1) W.sort = 0;
        if (tune_in_shoot && round == 2) W.sort = 2;
        W.sort_n = items_bound;
      }
      const int sort = W.sort;
      stamp();
      if (P.src.reemission_kind != REEMISSION_NONE && round > 0) {
        reemit_decide_kernel<<<decide_grid, 256, 0, s>>>(W);
        ++g_launches;
      }
      if (mode == ACC_HONLY) prepare_kernel<ACC_HONLY><<<prep_grid, 256, 0, s>>>(W);
      else prepare_kernel<ACC_FULL><<<prep_grid, 256, 0, s>>>(W);
      stamp();
      advance_after_prepare_kernel<<<1, 1, 0, s>>>(W.ctl, cap);
      if (sort == 1) {
        CUDA_OK(cudaMemsetAsync(W.hist, 0, (W.nbins + 1) * sizeof(uint32_t), s));
        sort_histogram_kernel<<<prep_grid, 256, 0, s>>>(W.ctl, W.key, W.hist);
        sort_scan_kernel<<<1, 1024, 0, s>>>(W.hist, W.nbins);
        sort_scatter_kernel<<<prep_grid, 256, 0, s>>>(W.ctl, W.key, W.hist, W.order);
        g_launches += 3;
      }
      if (sort == 2) {
        CUDA_OK(cub::DeviceRadixSort::SortPairs(ctx->sort_temp.p, sort_temp_bytes, W.key, ctx->sort_key_out.p,
                                                ctx->sort_iota.p, W.order, (int64_t)W.sort_n, 0, W.fine_key_bits + 1, s));
        g_launches += 2 + (W.fine_key_bits + 8) / 8; /* onesweep: histogram, scan, one pass per 8 key bits */
      }
      stamp();
      {
        const bool agg = (sort == 2 && W.agg);
        const unsigned march_grid = march_grids[agg ? 1 : 0];
        const bool prefetch = prefetch_cfg < 0 ? agg : (prefetch_cfg != 0);
        if (agg && mode == ACC_HONLY && lean_cfg) {
          /* march_coherent.cuh: the H-only coherent walk, variant by what the packets of this shoot can carry */
          using K = void (*)(const WavefrontParams);
          static const K variants[2][2][2][2] = {
#define CMIB_LEAN_V(H, PER, PR) {march_lean_kernel<H, PER, PR, 2>, march_lean_kernel<H, PER, PR, 3>}
              {{CMIB_LEAN_V(false, false, false), CMIB_LEAN_V(false, false, true)},
               {CMIB_LEAN_V(false, true, false), CMIB_LEAN_V(false, true, true)}},
              {{CMIB_LEAN_V(true, false, false), CMIB_LEAN_V(true, false, true)},
               {CMIB_LEAN_V(true, true, false), CMIB_LEAN_V(true, true, true)}}};
#undef CMIB_LEAN_V
          K k = variants[lean_heat][lean_periodic][prefetch ? 1 : 0][lean_steps == 2 ? 0 : 1];
          if (ctx->lean_grid[lean_heat][lean_periodic][prefetch ? 1 : 0][lean_steps == 2 ? 0 : 1] == 0) {
            CUDA_OK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_smem));
            int occ = 0;
            CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, MARCH_BLOCK, lean_smem));
            if (occ < 1) CMIB_FAIL("march_lean_kernel does not fit on an SM with %zu bytes of wall tables", lean_smem);
            ctx->lean_grid[lean_heat][lean_periodic][prefetch ? 1 : 0][lean_steps == 2 ? 0 : 1] = occ;
          }
          int bpm = ctx->lean_grid[lean_heat][lean_periodic][prefetch ? 1 : 0][lean_steps == 2 ? 0 : 1];
          if (const char *e = getenv("CMIB_MARCH_BLOCKS_PER_SM")) bpm = atoi(e) > 0 ? atoi(e) : bpm;
          k<<<(unsigned)(ctx->sm_count * bpm), MARCH_BLOCK, lean_smem, s>>>(W);
        } else
#define CMIB_LAUNCH_MARCH(M, A, R) march_kernel<M, A, R><<<march_grid, MARCH_BLOCK, 0, s>>>(W)
        if (mode == ACC_HONLY) {
          if (agg) { if (prefetch) CMIB_LAUNCH_MARCH(ACC_HONLY, true, true); else CMIB_LAUNCH_MARCH(ACC_HONLY, true, false); }
          else { if (prefetch) CMIB_LAUNCH_MARCH(ACC_HONLY, false, true); else CMIB_LAUNCH_MARCH(ACC_HONLY, false, false); }
        } else {
          if (agg) { if (prefetch) CMIB_LAUNCH_MARCH(ACC_FULL, true, true); else CMIB_LAUNCH_MARCH(ACC_FULL, true, false); }
          else { if (prefetch) CMIB_LAUNCH_MARCH(ACC_FULL, false, true); else CMIB_LAUNCH_MARCH(ACC_FULL, false, false); }
        }
#undef CMIB_LAUNCH_MARCH
      }
      stamp();
      advance_after_march_kernel<<<1, 1, 0, s>>>(W.ctl);
      g_launches += 4;
    }
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaMemcpyAsync(ctx->h_ctl, ctx->ctl.p, CTL_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    if (ctx->h_ctl[CTL_ERROR]) CMIB_FAIL("march kernel exceeded its pass limit (internal error)");
    bool done = false;
    for (int k = 0; k < group; ++k)
      if (ctx->h_ctl[CTL_STATUS + ((round - 1 - k) % CTL_STATUS_SLOTS)] == 0) done = true;
    if (done) break;
    if (tune_in_shoot && round == (uint64_t)group) { /* the stream is idle: rounds 1 and 2 are timed */
      float ms0 = 0.f, ms2 = 0.f;
      cudaEventElapsedTime(&ms0, ctx->tune_ev[0], ctx->tune_ev[1]);
      cudaEventElapsedTime(&ms2, ctx->tune_ev[1], ctx->tune_ev[2]);
      const double n0 = (double)ctx->h_ctl[CTL_STATUS + 1], n2 = (double)ctx->h_ctl[CTL_STATUS + 2];
      if (n0 > 0. && n2 > 0.) {
        ctx->tune_ns_per_crossing[0] = 1e6 * ms0 / n0; /* per queue entry here */
        ctx->tune_ns_per_crossing[2] = 1e6 * ms2 / n2;
        ctx->tuned_order = (ms2 / n2 < ms0 / n0) ? 2 : 0;
      }
      order_now = ctx->tuned_order;
    }
    if (ctx->h_ctl[CTL_REMAINING] == 0) {
      primaries_left = false;
      const uint64_t last = ctx->h_ctl[CTL_STATUS + ((round - 1) % CTL_STATUS_SLOTS)];
      items_bound = last < cap ? last : cap;
    }
  }
  ctx->shoot_rounds = round;
  if (P.hot_replicas > 0) {
    const int n = P.src.n_sources * HOT_CELLS * HOT_STRIDE;
    if (mode == ACC_HONLY) fold_hot_cells_kernel<ACC_HONLY><<<blocks_for(n, 128), 128, 0, s>>>(P);
    else fold_hot_cells_kernel<ACC_FULL><<<blocks_for(n, 128), 128, 0, s>>>(P);
    ++g_launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s));
  }
  if (tuning) {
    /* every group of rounds ends with a stream synchronisation: host time = device time here */
    const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - tune_t0).count();
    double crossings = 0.; /* counter 5 of the accumulator buffer: cell crossings (shoot.cuh) */
    CUDA_OK(cudaMemcpyAsync(&crossings, ctx->acc.p + 5, sizeof(double), cudaMemcpyDeviceToHost, s));
    CUDA_OK(cudaStreamSynchronize(s));
    crossings -= tune_crossings0;
    ctx->tune_ns_per_crossing[sort_cfg] = ns / (crossings > 1. ? crossings : 1.);
    if (sort_cfg == 2) ctx->tuned_order = (ctx->tune_ns_per_crossing[2] < ctx->tune_ns_per_crossing[0]) ? 2 : 0;
  }
  for (size_t k = 0; k + 3 < ev_used; k += 4) {
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, ctx->ev_pool[k], ctx->ev_pool[k + 1]);
    cudaEventElapsedTime(&b, ctx->ev_pool[k + 2], ctx->ev_pool[k + 3]);
    ctx->prepare_ms += a;
    ctx->march_ms += b;
  }
  return 0;
}

} // namespace

extern "C" {

int cmib_abi_version(void) { return CMIB_ABI_VERSION; }
const char *cmib_last_error(void) { return g_last_error.c_str(); }
void cmib_set_abort_on_error(int on) { g_abort_on_error = on ? 1 : 0; }
uint64_t cmib_kernel_launch_count(void) { return g_launches.load(); }

int cmib_create(const cmib_grid_desc *grid, int device, cmib_context **out) {
  if (!grid || !out) CMIB_FAIL("null argument");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    CMIB_FAIL("no CUDA device available (%s); this library has no CPU path",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
  if (device < 0 || device >= ndev) CMIB_FAIL("device %d out of range [0,%d)", device, ndev);
  for (int d = 0; d < 3; ++d) {
    if (grid->ncell[d] <= 0) CMIB_FAIL("number of cells must be positive");
    if (!(grid->sides[d] > 0.)) CMIB_FAIL("box sides must be positive");
  }
  CUDA_OK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10) CMIB_FAIL("device %d is sm_%d%d; this library is built for sm_100a only", device,
                                 prop.major, prop.minor);
  cmib_context *ctx = new cmib_context();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  CUDA_OK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
  GridGeom &g = ctx->geom;
  for (int d = 0; d < 3; ++d) {
    g.anchor[d] = grid->anchor[d];
    g.sides[d] = grid->sides[d];
    g.ncell[d] = grid->ncell[d];
    g.periodic[d] = grid->periodic[d] ? 1 : 0;
    /* CartesianDensityGrid.cpp:80-86 */
    g.cellside[d] = grid->sides[d] / grid->ncell[d];
    g.inv_cellside[d] = 1. / g.cellside[d];
  }
  g.cell_volume = g.cellside[0] * g.cellside[1] * g.cellside[2];
  g.ncells = (int64_t)g.ncell[0] * g.ncell[1] * g.ncell[2];
  const size_t nc = (size_t)g.ncells;
  CUDA_OK(ctx->cells.resize(nc));
  CUDA_OK(ctx->cells_h.resize(nc));
  CUDA_OK(cudaMemsetAsync(ctx->cells_h.p, 0, nc * sizeof(double2), ctx->stream));
  CUDA_OK(ctx->xmetal.resize(nc * 12));
  CUDA_OK(ctx->heat_norm.resize(nc * 2));
  CUDA_OK(cudaMemsetAsync(ctx->cells.p, 0, nc * sizeof(CellOpacity), ctx->stream));
  CUDA_OK(cudaMemsetAsync(ctx->xmetal.p, 0, nc * 12 * sizeof(double), ctx->stream));
  CUDA_OK(cudaMemsetAsync(ctx->heat_norm.p, 0, nc * 2 * sizeof(double), ctx->stream));
  memset(&ctx->src, 0, sizeof(ctx->src));
  ctx->src.discrete_weight = 1.;
  ctx->src.xs_kind = XS_VERNER;
  ctx->src.spectrum.kind = SPECTRUM_MONOCHROMATIC;
  ctx->src.spectrum.mono_frequency = ev_to_hz(13.6);
  ctx->src.reemission_kind = REEMISSION_NONE;
  memset(&ctx->rr, 0, sizeof(ctx->rr));
  ctx->rr.kind = RR_VERNER;
  ctx->tp.do_temperature = 0;
  ctx->tp.min_iterations = 3;
  ctx->tp.epsilon = 1.e-3;
  ctx->tp.max_iterations = 100;
  ctx->tp.pahfac = 0.;
  ctx->tp.crfac = 0.;
  ctx->tp.crlim = 0.75;
  ctx->tp.crscale = 1.33333 * 3.086e19;
  ctx->tp.min_ionized_T = 4000.;
  /* DensityGrid.hpp:219-222 */
  ctx->nu_H = ev_to_hz(13.6);
  ctx->nu_He = ev_to_hz(24.6);
  ctx->acc_mode = -1;
  if (ensure_acc(ctx)) {
    delete ctx;
    return 1;
  }
  *out = ctx;
  return 0;
}

int cmib_destroy(cmib_context *ctx) {
  if (!ctx) return 0;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  cudaStreamDestroy(ctx->stream);
  if (ctx->h_ctl) cudaFreeHost(ctx->h_ctl);
  for (cudaEvent_t e : ctx->ev_pool) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->tune_ev)
    if (e) cudaEventDestroy(e);
  delete ctx;
  return 0;
}

int cmib_synchronize(cmib_context *ctx) {
  CHECK_CTX(ctx);
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cmib_stream(cmib_context *ctx, void **stream) {
  CHECK_CTX(ctx);
  *stream = (void *)ctx->stream;
  return 0;
}

/* ---- grid state ------------------------------------------------------------ */
int cmib_upload_cells(cmib_context *ctx, const double *n, const double *T, const double *x,
                      const double *cr_factor) {
  CHECK_CTX(ctx);
  if (!n || !T || !x) CMIB_FAIL("null cell array");
  const size_t nc = (size_t)ctx->geom.ncells;
  CUDA_OK(ctx->stage.resize(nc * 18));
  double *s = ctx->stage.p;
  CUDA_OK(cudaMemcpyAsync(s, n, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(s + nc, T, nc * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(s + 2 * nc, x, nc * 14 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  pack_cells_kernel<<<blocks_for(nc, 256), 256, 0, ctx->stream>>>((int64_t)nc, s, s + nc, s + 2 * nc,
                                                                  ctx->cells.p, ctx->cells_h.p, ctx->xmetal.p);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  if (cr_factor) {
    CUDA_OK(ctx->cr_factor.upload(cr_factor, nc, ctx->stream));
    ctx->have_cr_factor = true;
  } else {
    ctx->have_cr_factor = false;
  }
  ctx->reemit_prob_valid = false;
  CUDA_OK(cudaStreamSynchronize(ctx->stream)); /* host arrays may be reused on return */
  return 0;
}

int cmib_download_cells(cmib_context *ctx, double *n, double *T, double *x, double *heat) {
  CHECK_CTX(ctx);
  const size_t nc = (size_t)ctx->geom.ncells;
  CUDA_OK(ctx->stage.resize(nc * 18));
  double *s = ctx->stage.p;
  unpack_cells_kernel<<<blocks_for(nc, 256), 256, 0, ctx->stream>>>(
      (int64_t)nc, ctx->cells.p, ctx->xmetal.p, ctx->heat_norm.p, s, s + nc, s + 2 * nc, s + 16 * nc);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  if (n) CUDA_OK(cudaMemcpyAsync(n, s, nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (T) CUDA_OK(cudaMemcpyAsync(T, s + nc, nc * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (x) CUDA_OK(cudaMemcpyAsync(x, s + 2 * nc, nc * 14 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (heat) CUDA_OK(cudaMemcpyAsync(heat, s + 16 * nc, nc * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cmib_download_accumulators(cmib_context *ctx, double *J, double *heat) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  const size_t nc = (size_t)ctx->geom.ncells;
  CUDA_OK(ctx->stage.resize(nc * 18));
  double *s = ctx->stage.p;
  if (ctx->acc_mode == ACC_HONLY)
    unpack_acc_kernel<ACC_HONLY><<<blocks_for(nc, 256), 256, 0, ctx->stream>>>(
        (int64_t)nc, ctx->acc.p, ctx->honly_cell_stride(), ctx->honly_term_stride(), ctx->honly_offset(), s, s + 14 * nc);
  else
    unpack_acc_kernel<ACC_FULL><<<blocks_for(nc, 256), 256, 0, ctx->stream>>>((int64_t)nc, ctx->acc.p, 0, 0, 0, s, s + 14 * nc);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  if (J) CUDA_OK(cudaMemcpyAsync(J, s, nc * 14 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (heat) CUDA_OK(cudaMemcpyAsync(heat, s + 14 * nc, nc * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cmib_reset_accumulators(cmib_context *ctx) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  CUDA_OK(cudaMemsetAsync(ctx->acc.p, 0, ctx->acc.n * sizeof(double), ctx->stream));
  return 0;
}

/* ---- plugins --------------------------------------------------------------- */
int cmib_set_abundances(cmib_context *ctx, const double *abundances) {
  CHECK_CTX(ctx);
  if (!abundances) CMIB_FAIL("null abundances");
  for (int k = 0; k < NUM_ELEMENTS; ++k) ctx->abund[k] = abundances[k];
  ctx->src.A_He = abundances[EL_He];
  return 0;
}

int cmib_set_bimodal_cross_sections(cmib_context *ctx, double frequency_limit, const double *low, const double *high) {
  CHECK_CTX(ctx);
  if (!low || !high) CMIB_FAIL("Bimodal cross sections need 14 values below and 14 above the limit");
  ctx->src.xs_kind = XS_BIMODAL;
  ctx->src.xs_limit = frequency_limit;
  for (int k = 0; k < NUM_IONS; ++k) {
    ctx->src.xs_fixed[k] = low[k];
    ctx->src.xs_high[k] = high[k];
  }
  return ensure_acc(ctx);
}

int cmib_set_cross_sections(cmib_context *ctx, int kind, const double *fixed) {
  CHECK_CTX(ctx);
  if (kind == CMIB_CROSS_SECTIONS_VERNER) {
    ctx->src.xs_kind = XS_VERNER;
  } else if (kind == CMIB_CROSS_SECTIONS_FIXED_VALUE) {
    if (!fixed) CMIB_FAIL("FixedValue cross sections need 14 values");
    ctx->src.xs_kind = XS_FIXED;
    for (int k = 0; k < NUM_IONS; ++k) ctx->src.xs_fixed[k] = fixed[k];
  } else {
    CMIB_FAIL("Unknown CrossSections type: %d", kind);
  }
  return ensure_acc(ctx);
}

int cmib_set_recombination_rates(cmib_context *ctx, int kind, const double *fixed) {
  CHECK_CTX(ctx);
  if (kind == CMIB_RECOMBINATION_VERNER) {
    ctx->rr.kind = RR_VERNER;
  } else if (kind == CMIB_RECOMBINATION_FIXED_VALUE) {
    if (!fixed) CMIB_FAIL("FixedValue recombination rates need 14 values");
    ctx->rr.kind = RR_FIXED;
    for (int k = 0; k < NUM_IONS; ++k) ctx->rr.fixed[k] = fixed[k];
  } else {
    CMIB_FAIL("Unknown RecombinationRates type: %d", kind);
  }
  return 0;
}

int cmib_set_sources(cmib_context *ctx, int32_t n, const double *positions, const double *weights,
                     double total_luminosity) {
  CHECK_CTX(ctx);
  if (n == 0) { /* PhotonSourceDistribution: None — only legal together with a continuous source */
    ctx->src.n_sources = 0;
    ctx->src.src_pos = nullptr;
    ctx->src.src_cum = nullptr;
    ctx->hot_replicas = 0;
    ctx->discrete_luminosity = 0.;
    ctx->update_source_weights();
    return 0;
  }
  if (n < 0 || !positions || !weights) CMIB_FAIL("need at least one discrete source");
  /* PhotonSource.cpp:74-100 */
  std::vector<double> cum(n);
  for (int i = 0; i < n; ++i) cum[i] = (i > 0 ? cum[i - 1] : 0.) + weights[i];
  if (std::abs(cum[n - 1] - 1.) > 1.e-9) CMIB_FAIL("Discrete source weights do not sum to 1.0 (%g)!", cum[n - 1]);
  cum[n - 1] = 1.;
  CUDA_OK(ctx->d_src_pos.upload(positions, (size_t)n * 3, ctx->stream));
  CUDA_OK(ctx->d_src_cum.upload(cum.data(), (size_t)n, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  /* cell of every source, as get_cell_indices finds it (CartesianDensityGrid.cpp:152-161) */
  ctx->h_src_cell.assign(n, 0u);
  bool hot_ok = (n <= HOT_MAX_SOURCES);
  for (int i = 0; i < n && hot_ok; ++i) {
    uint32_t packed = 0;
    for (int d = 0; d < 3; ++d) {
      const double v = (positions[3 * i + d] - ctx->geom.anchor[d]) * ctx->geom.inv_cellside[d];
      const long long idx = (v != v) ? -1 : (long long)v;
      if (idx < 0 || idx >= ctx->geom.ncell[d] || idx > 1022) hot_ok = false;
      else packed |= (uint32_t)idx << (10 * d);
    }
    ctx->h_src_cell[i] = packed;
  }
  if (const char *e = getenv("CMIB_HOT_REPLICAS")) ctx->hot_replicas = atoi(e); else ctx->hot_replicas = 64;
  if (!hot_ok || ctx->geom.periodic[0] || ctx->geom.periodic[1] || ctx->geom.periodic[2]) ctx->hot_replicas = 0;
  if (ctx->hot_replicas > 0) {
    CUDA_OK(ctx->d_src_cell.upload(ctx->h_src_cell.data(), (size_t)n, ctx->stream));
    const size_t nd = (size_t)ctx->hot_replicas * HOT_MAX_SOURCES * HOT_CELLS * HOT_STRIDE;
    if (ctx->hot_acc.n != nd) {
      CUDA_OK(ctx->hot_acc.resize(nd));
      CUDA_OK(cudaMemsetAsync(ctx->hot_acc.p, 0, nd * sizeof(double), ctx->stream));
    }
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  ctx->src.n_sources = n;
  ctx->src.src_pos = ctx->d_src_pos.p;
  ctx->src.src_cum = ctx->d_src_cum.p;
  ctx->discrete_luminosity = total_luminosity;
  ctx->update_source_weights();
  return 0;
}

int cmib_set_distant_star_position(cmib_context *ctx, const double position[3]) {
  CHECK_CTX(ctx);
  if (!position) CMIB_FAIL("null argument");
  int num_exposed = 0;
  for (int d = 0; d < 3; ++d) {
    const double bottom = ctx->geom.anchor[d], top = ctx->geom.anchor[d] + ctx->geom.sides[d];
    ctx->src.star_position[d] = position[d];
    ctx->src.star_exposed[d] = (position[d] < bottom) ? -1 : ((position[d] > top) ? 1 : 0);
    num_exposed += (ctx->src.star_exposed[d] != 0);
  }
  if (num_exposed == 0) CMIB_FAIL("External stellar source lies inside the simulation box. This will not work!");
  ctx->star_position_set = true;
  return 0;
}

int cmib_set_planar_source_geometry(cmib_context *ctx, int normal_axis, double intercept, const double anchor[2],
                                    const double sides[2]) {
  CHECK_CTX(ctx);
  if (normal_axis < 0 || normal_axis > 2 || !anchor || !sides) CMIB_FAIL("normal axis must be 0, 1 or 2");
  ctx->src.planar_axis = normal_axis;
  ctx->src.planar_intercept = intercept;
  for (int k = 0; k < 2; ++k) {
    ctx->src.planar_anchor[k] = anchor[k];
    ctx->src.planar_sides[k] = sides[k];
  }
  ctx->planar_geometry_set = true;
  ctx->disc_geometry_set = false; /* the two geometries share the axis and the intercept */
  return 0;
}

int cmib_set_extended_disc_geometry(cmib_context *ctx, int normal_axis, double origin, double scale_height) {
  CHECK_CTX(ctx);
  if (normal_axis < 0 || normal_axis > 2) CMIB_FAIL("normal axis must be 0, 1 or 2");
  if (!(scale_height > 0.)) CMIB_FAIL("the scale height of the disc must be positive");
  /* a disc whose Gaussian never reaches the box would redraw forever: ask for 10 sigma at most */
  const double bottom = ctx->geom.anchor[normal_axis], top = bottom + ctx->geom.sides[normal_axis];
  if (origin < bottom - 10. * scale_height || origin > top + 10. * scale_height)
    CMIB_FAIL("the disc lies more than 10 scale heights outside the simulation box");
  ctx->src.planar_axis = normal_axis;
  ctx->src.planar_intercept = origin;
  ctx->src.disc_scale_height = scale_height;
  ctx->disc_geometry_set = true;
  ctx->planar_geometry_set = false; /* the two geometries share the axis and the intercept */
  return 0;
}

int cmib_set_spiral_galaxy_geometry(cmib_context *ctx, double scale_length_stars, double scale_height_stars,
                                    double bulge_over_total_ratio) {
  CHECK_CTX(ctx);
  if (!(scale_length_stars > 0.) || !(scale_height_stars > 0.)) CMIB_FAIL("the scale length and height of the stellar disc must be positive");
  if (!(bulge_over_total_ratio >= 0.) || bulge_over_total_ratio > 1.) CMIB_FAIL("the bulge over total ratio must lie in [0, 1]");
  /* the galaxy sits at the origin: a box that does not hold the origin would reject (nearly) every position */
  for (int d = 0; d < 3; ++d)
    if (ctx->geom.anchor[d] > 0. || ctx->geom.anchor[d] + ctx->geom.sides[d] <= 0.)
      CMIB_FAIL("the SpiralGalaxy source is centred on the origin, which lies outside the simulation box");
  std::vector<double> tables(2 * (GALAXY_NBIN + 1));
  build_galaxy_model(ctx->geom.anchor, scale_length_stars, scale_height_stars, bulge_over_total_ratio, ctx->src.galaxy,
                     tables.data(), tables.data() + GALAXY_NBIN + 1);
  CUDA_OK(ctx->d_galaxy_tables.upload(tables.data(), tables.size(), ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  ctx->src.galaxy_w = ctx->d_galaxy_tables.p;
  ctx->src.galaxy_cdf = ctx->d_galaxy_tables.p + GALAXY_NBIN + 1;
  ctx->galaxy_geometry_set = true;
  return 0;
}

int cmib_set_continuous_source(cmib_context *ctx, int kind, double luminosity, int spectrum_kind,
                               double spectrum_param) {
  CHECK_CTX(ctx);
  if (kind == CMIB_CONTINUOUS_NONE) {
    ctx->src.continuous_kind = CONTINUOUS_NONE;
    ctx->continuous_luminosity = 0.;
    ctx->update_source_weights();
    return 0;
  }
  if (kind != CMIB_CONTINUOUS_ISOTROPIC && kind != CMIB_CONTINUOUS_PLANAR && kind != CMIB_CONTINUOUS_DISTANT_STAR &&
      kind != CMIB_CONTINUOUS_EXTENDED_DISC && kind != CMIB_CONTINUOUS_SPIRAL_GALAXY)
    CMIB_FAIL("Unknown ContinuousPhotonSource type: %d", kind);
  if (kind == CMIB_CONTINUOUS_SPIRAL_GALAXY && !ctx->galaxy_geometry_set)
    CMIB_FAIL("call cmib_set_spiral_galaxy_geometry before selecting the SpiralGalaxy continuous source");
  if (kind == CMIB_CONTINUOUS_EXTENDED_DISC && !ctx->disc_geometry_set)
    CMIB_FAIL("call cmib_set_extended_disc_geometry before selecting the ExtendedDisc continuous source");
  if (kind == CMIB_CONTINUOUS_DISTANT_STAR && !ctx->star_position_set)
    CMIB_FAIL("call cmib_set_distant_star_position before selecting the DistantStar continuous source");
  if (kind == CMIB_CONTINUOUS_PLANAR && !ctx->planar_geometry_set)
    CMIB_FAIL("call cmib_set_planar_source_geometry before selecting the Planar continuous source");
  if (!(luminosity > 0.)) CMIB_FAIL("the continuous source needs a positive luminosity (surface area x total flux)");
  /* CMIB_SPECTRUM_TABULATED: the table was (or will be) given with cmib_set_spectrum_table(ctx, 1, ...) */
  if (spectrum_kind != CMIB_SPECTRUM_TABULATED &&
      set_spectrum_model(ctx, ctx->src.cont_spectrum, ctx->h_cont_planck, ctx->d_cont_planck, ctx->d_cont_planck_guide,
                         spectrum_kind, spectrum_param))
    return 1;
  ctx->src.continuous_kind = (kind == CMIB_CONTINUOUS_PLANAR) ? CONTINUOUS_PLANAR
                             : (kind == CMIB_CONTINUOUS_DISTANT_STAR)
                                   ? CONTINUOUS_DISTANT_STAR
                                   : (kind == CMIB_CONTINUOUS_EXTENDED_DISC)
                                         ? CONTINUOUS_EXTENDED_DISC
                                         : (kind == CMIB_CONTINUOUS_SPIRAL_GALAXY ? CONTINUOUS_SPIRAL_GALAXY : CONTINUOUS_ISOTROPIC);
  ctx->continuous_luminosity = luminosity;
  ctx->update_source_weights();
  return 0;
}

int cmib_set_spectrum(cmib_context *ctx, int kind, double param) {
  CHECK_CTX(ctx);
  return set_spectrum_model(ctx, ctx->src.spectrum, ctx->h_planck, ctx->d_planck, ctx->d_planck_guide, kind, param);
}

int cmib_set_spectrum_table(cmib_context *ctx, int role, int32_t n, const double *frequencies,
                            const double *cumulative_distribution) {
  CHECK_CTX(ctx);
  if (role != 0 && role != 1) CMIB_FAIL("role must be 0 (PhotonSourceSpectrum) or 1 (ContinuousPhotonSourceSpectrum)");
  if (n < 2 || !frequencies || !cumulative_distribution) CMIB_FAIL("a tabulated spectrum needs at least two frequencies");
  for (int32_t i = 1; i < n; ++i)
    if (cumulative_distribution[i] < cumulative_distribution[i - 1])
      CMIB_FAIL("the cumulative distribution of a tabulated spectrum must not decrease (entry %d)", (int)i);
  CUDA_OK(ctx->d_spec_freq[role].upload(frequencies, (size_t)n, ctx->stream));
  CUDA_OK(ctx->d_spec_cdf[role].upload(cumulative_distribution, (size_t)n, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  SpectrumModel &sp = role ? ctx->src.cont_spectrum : ctx->src.spectrum;
  sp.kind = SPECTRUM_TABULATED;
  sp.n = n;
  sp.freq = ctx->d_spec_freq[role].p;
  sp.cdf = ctx->d_spec_cdf[role].p;
  return 0;
}

int cmib_set_reemission(cmib_context *ctx, int kind, double probability, double frequency) {
  CHECK_CTX(ctx);
  if (kind == CMIB_REEMISSION_NONE) {
    ctx->src.reemission_kind = REEMISSION_NONE;
  } else if (kind == CMIB_REEMISSION_FIXED_VALUE) {
    ctx->src.reemission_kind = REEMISSION_FIXED;
    ctx->src.fixed_reemission_probability = probability;
    ctx->src.fixed_reemission_frequency = frequency;
  } else if (kind == CMIB_REEMISSION_PHYSICAL) {
    const SourceModel m = ctx->src;
    auto sigma_of = [m](int ion) {
      return [m, ion](double nu) {
        if (m.xs_kind == XS_VERNER) return verner_cross_section(ion, nu);
        if (m.xs_kind == XS_BIMODAL) return nu < m.xs_limit ? m.xs_fixed[ion] : m.xs_high[ion];
        return m.xs_fixed[ion];
      };
    };
    host::build_lyc_table(0, sigma_of(ION_H_n), ctx->h_hlyc_freq, ctx->h_hlyc_temp, ctx->h_hlyc_cdf);
    host::build_lyc_table(1, sigma_of(ION_He_n), ctx->h_helyc_freq, ctx->h_helyc_temp, ctx->h_helyc_cdf);
    host::build_he2pc_table(ctx->h_he2pc_freq, ctx->h_he2pc_cdf);
    CUDA_OK(ctx->d_hlyc_freq.upload(ctx->h_hlyc_freq.data(), ctx->h_hlyc_freq.size(), ctx->stream));
    CUDA_OK(ctx->d_hlyc_temp.upload(ctx->h_hlyc_temp.data(), ctx->h_hlyc_temp.size(), ctx->stream));
    CUDA_OK(ctx->d_hlyc_cdf.upload(ctx->h_hlyc_cdf.data(), ctx->h_hlyc_cdf.size(), ctx->stream));
    CUDA_OK(ctx->d_helyc_freq.upload(ctx->h_helyc_freq.data(), ctx->h_helyc_freq.size(), ctx->stream));
    CUDA_OK(ctx->d_helyc_temp.upload(ctx->h_helyc_temp.data(), ctx->h_helyc_temp.size(), ctx->stream));
    CUDA_OK(ctx->d_helyc_cdf.upload(ctx->h_helyc_cdf.data(), ctx->h_helyc_cdf.size(), ctx->stream));
    CUDA_OK(ctx->d_he2pc_freq.upload(ctx->h_he2pc_freq.data(), ctx->h_he2pc_freq.size(), ctx->stream));
    CUDA_OK(ctx->d_he2pc_cdf.upload(ctx->h_he2pc_cdf.data(), ctx->h_he2pc_cdf.size(), ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    ctx->src.hlyc_freq = ctx->d_hlyc_freq.p;
    ctx->src.hlyc_temp = ctx->d_hlyc_temp.p;
    ctx->src.hlyc_cdf = ctx->d_hlyc_cdf.p;
    ctx->src.helyc_freq = ctx->d_helyc_freq.p;
    ctx->src.helyc_temp = ctx->d_helyc_temp.p;
    ctx->src.helyc_cdf = ctx->d_helyc_cdf.p;
    ctx->src.he2pc_freq = ctx->d_he2pc_freq.p;
    ctx->src.he2pc_cdf = ctx->d_he2pc_cdf.p;
    if (make_guides(ctx, ctx->h_hlyc_cdf.data(), LYC_NUMTEMP, ctx->d_hlyc_guide, &ctx->src.hlyc_guide)) return 1;
    if (make_guides(ctx, ctx->h_helyc_cdf.data(), LYC_NUMTEMP, ctx->d_helyc_guide, &ctx->src.helyc_guide)) return 1;
    if (make_guides(ctx, ctx->h_he2pc_cdf.data(), 1, ctx->d_he2pc_guide, &ctx->src.he2pc_guide)) return 1;
    ctx->src.reemission_kind = REEMISSION_PHYSICAL;
    CUDA_OK(ctx->reemit_prob.resize((size_t)ctx->geom.ncells * NUM_REEMIT));
    ctx->reemit_prob_valid = false;
  } else {
    CMIB_FAIL("Unknown DiffuseReemissionHandler type: %d", kind);
  }
  return 0;
}

int cmib_set_temperature_params(cmib_context *ctx, const cmib_temperature_params *p) {
  CHECK_CTX(ctx);
  if (!p) CMIB_FAIL("null parameters");
  ctx->tp.do_temperature = p->do_temperature_calculation;
  ctx->tp.min_iterations = p->minimum_number_of_iterations;
  ctx->tp.epsilon = p->epsilon_convergence;
  ctx->tp.max_iterations = p->maximum_number_of_iterations;
  ctx->tp.pahfac = p->pah_heating_factor;
  ctx->tp.crfac = p->cosmic_ray_heating_factor;
  ctx->tp.crlim = p->cosmic_ray_heating_limit;
  ctx->tp.crscale = p->cosmic_ray_heating_scale_length;
  ctx->tp.min_ionized_T = p->minimum_ionized_temperature;
  return 0;
}

/* ---- iteration ------------------------------------------------------------- */
int cmib_update_reemission_probabilities(cmib_context *ctx) {
  CHECK_CTX(ctx);
  if (ctx->src.reemission_kind != REEMISSION_PHYSICAL) return 0;
  const int64_t nc = ctx->geom.ncells;
  reemission_probabilities_kernel<<<blocks_for(nc, 256), 256, 0, ctx->stream>>>(nc, ctx->cells.p,
                                                                                 ctx->reemit_prob.p);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  ctx->reemit_prob_valid = true;
  return 0;
}

int cmib_shoot(cmib_context *ctx, uint64_t n_packets, uint64_t packet_offset, uint64_t seed,
               uint32_t iteration, double *totweight, double *typecount) {
  CHECK_CTX(ctx);
  if (ctx->src.n_sources <= 0 && ctx->src.continuous_kind == CONTINUOUS_NONE) CMIB_FAIL("no photon sources set");
  if (ensure_acc(ctx)) return 1;
  if (ctx->src.reemission_kind == REEMISSION_PHYSICAL && !ctx->reemit_prob_valid)
    if (cmib_update_reemission_probabilities(ctx)) return 1;
  double before[5] = {0., 0., 0., 0., 0.};
  if (totweight || typecount) {
    CUDA_OK(cudaMemcpyAsync(before, ctx->acc.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  if (n_packets > 0) {
    ShootParams P;
    P.geom = ctx->geom;
    P.src = ctx->src;
    P.cells = ctx->cells.p;
    P.cells_h = ctx->cells_h.p;
    P.reemit_prob = ctx->reemit_prob.p;
    P.acc = ctx->acc.p;
    P.honly_cell_stride = ctx->honly_cell_stride();
    P.honly_term_stride = ctx->honly_term_stride();
    P.honly_offset = ctx->honly_offset();
    P.hot_acc = ctx->hot_acc.p;
    P.src_cell = ctx->d_src_cell.p;
    P.hot_replicas = ctx->hot_replicas;
    P.nu_H = ctx->nu_H;
    P.nu_He = ctx->nu_He;
    P.seed = seed;
    P.iteration = iteration;
    P.packet_offset = packet_offset;
    P.n_packets = n_packets;
    if (ctx->shoot_algorithm == 1) {
      const int bs = 256;
      /* persistent-style grid: a multiple of the SM count, packets are strided over it */
      uint64_t want = (n_packets + bs - 1) / bs;
      uint64_t cap = (uint64_t)ctx->sm_count * 8;
      unsigned grid = (unsigned)(want < cap ? want : cap);
      if (ctx->acc_mode == ACC_HONLY)
        shoot_kernel<ACC_HONLY><<<grid, bs, 0, ctx->stream>>>(P);
      else
        shoot_kernel<ACC_FULL><<<grid, bs, 0, ctx->stream>>>(P);
      ++g_launches;
      CUDA_OK(cudaGetLastError());
    } else {
      if (shoot_wavefront(ctx, P)) return 1;
    }
  }
  if (totweight || typecount) {
    double after[5];
    CUDA_OK(cudaMemcpyAsync(after, ctx->acc.p, 5 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    if (totweight) *totweight = after[0] - before[0];
    if (typecount)
      for (int t = 0; t < 4; ++t) typecount[t] = after[1 + t] - before[1 + t];
  }
  return 0;
}

int cmib_update_state(cmib_context *ctx, uint32_t loop, double totweight) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  UpdateParams P;
  P.geom = ctx->geom;
  P.cells = ctx->cells.p;
  P.cells_h = ctx->cells_h.p;
  P.xmetal = ctx->xmetal.p;
  P.heat_norm = ctx->heat_norm.p;
  P.cr_factor = ctx->have_cr_factor ? ctx->cr_factor.p : nullptr;
  P.acc = ctx->acc.p;
  P.honly_cell_stride = ctx->honly_cell_stride();
  P.honly_term_stride = ctx->honly_term_stride();
  P.honly_offset = ctx->honly_offset();
  P.luminosity = ctx->luminosity;
  P.totweight = totweight;
  for (int k = 0; k < NUM_ELEMENTS; ++k) P.abund[k] = ctx->abund[k];
  P.rr = ctx->rr;
  P.tp = ctx->tp;
  /* TemperatureCalculator.cpp:948: strictly greater */
  P.solve_temperature = (ctx->tp.do_temperature && loop > ctx->tp.min_iterations) ? 1 : 0;
  const int64_t nc = ctx->geom.ncells;
  const char *simple = getenv("CMIB_UPDATE_SIMPLE");
  if (P.solve_temperature && !(simple && simple[0] == '1')) {
    /* temperature solve: persistent warps with dynamic cell hand-out (kernels.cuh) */
    if (!ctx->upd_counter.p) {
      CUDA_OK(ctx->upd_counter.resize(1));
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->update_blocks_per_sm[ACC_FULL],
                                                            update_temperature_kernel<ACC_FULL>, 128, 0));
      CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->update_blocks_per_sm[ACC_HONLY],
                                                            update_temperature_kernel<ACC_HONLY>, 128, 0));
    }
    CUDA_OK(cudaMemsetAsync(ctx->upd_counter.p, 0, sizeof(unsigned long long), ctx->stream));
    int bpm = ctx->update_blocks_per_sm[ctx->acc_mode];
    if (bpm < 1) bpm = 1;
    unsigned grid = (unsigned)(ctx->sm_count * bpm);
    const unsigned need = blocks_for(nc, 128);
    if (grid > need) grid = need;
    if (ctx->acc_mode == ACC_HONLY)
      update_temperature_kernel<ACC_HONLY><<<grid, 128, 0, ctx->stream>>>(P, ctx->upd_counter.p);
    else
      update_temperature_kernel<ACC_FULL><<<grid, 128, 0, ctx->stream>>>(P, ctx->upd_counter.p);
  } else if (ctx->acc_mode == ACC_HONLY) {
    update_state_kernel<ACC_HONLY><<<blocks_for(nc, 128), 128, 0, ctx->stream>>>(P);
  } else {
    update_state_kernel<ACC_FULL><<<blocks_for(nc, 128), 128, 0, ctx->stream>>>(P);
  }
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  ctx->reemit_prob_valid = false;
  return 0;
}

int cmib_set_shoot_algorithm(cmib_context *ctx, int algorithm) {
  CHECK_CTX(ctx);
  if (algorithm != 0 && algorithm != 1) CMIB_FAIL("unknown shoot algorithm %d", algorithm);
  ctx->shoot_algorithm = algorithm;
  return 0;
}

int cmib_set_shoot_timing(cmib_context *ctx, int on) {
  CHECK_CTX(ctx);
  ctx->timing = on != 0;
  return 0;
}

int cmib_shoot_timing(cmib_context *ctx, double *prepare_ms, double *march_ms, uint64_t *rounds,
                      double *accumulator_adds) {
  CHECK_CTX(ctx);
  if (prepare_ms) *prepare_ms = ctx->prepare_ms;
  if (march_ms) *march_ms = ctx->march_ms;
  if (rounds) *rounds = ctx->shoot_rounds;
  if (accumulator_adds) {
    double v = 0.;
    CUDA_OK(cudaMemcpyAsync(&v, ctx->acc.p + 7, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    *accumulator_adds = v;
  }
  return 0;
}

int cmib_shoot_statistics(cmib_context *ctx, double *cell_crossings, double *emissions) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  double c[ACC_COUNTERS];
  CUDA_OK(cudaMemcpyAsync(c, ctx->acc.p, ACC_COUNTERS * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  if (cell_crossings) *cell_crossings = c[5];
  if (emissions) *emissions = c[6];
  return 0;
}

int cmib_shoot_optical_depth(cmib_context *ctx, double *tau_traversed) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  if (!tau_traversed) CMIB_FAIL("null argument");
  CUDA_OK(cudaMemcpyAsync(tau_traversed, ctx->acc.p + 8, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cmib_accumulator_buffer(cmib_context *ctx, void **device_ptr, uint64_t *n_doubles) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  if (device_ptr) *device_ptr = ctx->acc.p;
  if (n_doubles) *n_doubles = ctx->acc.n;
  return 0;
}

/* ---- test hooks ------------------------------------------------------------ */

int cmib_march_packets(cmib_context *ctx, int64_t np, const double *pos, const double *dir,
                       const double *sigma, const double *sigma_He_corr, const double *nu,
                       const double *weight, const double *tau, double *final_pos,
                       int64_t *final_cell, int32_t *nsteps, int32_t max_trace, int64_t *trace) {
  CHECK_CTX(ctx);
  if (np <= 0) return 0;
  if (!ctx->force_full) {
    ctx->force_full = true; /* explicit packets carry all 14 cross sections */
  }
  if (ensure_acc(ctx)) return 1;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  MarchPacketsParams P;
  P.geom = ctx->geom;
  P.cells = ctx->cells.p;
  P.acc = ctx->acc.p;
  P.nu_H = ctx->nu_H;
  P.nu_He = ctx->nu_He;
  P.np = np;
  double *d_pos, *d_dir, *d_sigma, *d_she, *d_nu, *d_w, *d_tau, *d_fp;
  int64_t *d_fc, *d_trace = nullptr;
  int32_t *d_ns;
  CUDA_OK(sc.in(&d_pos, pos, (size_t)np * 3, s));
  CUDA_OK(sc.in(&d_dir, dir, (size_t)np * 3, s));
  CUDA_OK(sc.in(&d_sigma, sigma, (size_t)np * NUM_IONS, s));
  CUDA_OK(sc.in(&d_she, sigma_He_corr, (size_t)np, s));
  CUDA_OK(sc.in(&d_nu, nu, (size_t)np, s));
  CUDA_OK(sc.in(&d_w, weight, (size_t)np, s));
  CUDA_OK(sc.in(&d_tau, tau, (size_t)np, s));
  CUDA_OK(sc.out(&d_fp, (size_t)np * 3));
  CUDA_OK(sc.out(&d_fc, (size_t)np));
  CUDA_OK(sc.out(&d_ns, (size_t)np));
  if (trace && max_trace > 0) CUDA_OK(sc.out(&d_trace, (size_t)np * max_trace));
  P.pos = d_pos; P.dir = d_dir; P.sigma = d_sigma; P.sigma_He_corr = d_she; P.nu = d_nu;
  P.weight = d_w; P.tau = d_tau; P.final_pos = d_fp; P.final_cell = d_fc; P.nsteps = d_ns;
  P.max_trace = d_trace ? max_trace : 0;
  P.trace = d_trace;
  march_packets_kernel<<<blocks_for(np, 128), 128, 0, s>>>(P);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(final_pos, d_fp, (size_t)np * 3, s));
  CUDA_OK(sc.back(final_cell, d_fc, (size_t)np, s));
  CUDA_OK(sc.back(nsteps, d_ns, (size_t)np, s));
  if (d_trace) CUDA_OK(sc.back(trace, d_trace, (size_t)np * max_trace, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_integrate_optical_depth(cmib_context *ctx, int64_t np, const double *pos, const double *dir,
                                 const double *sigma_H, const double *sigma_He_corr, double *optical_depth) {
  CHECK_CTX(ctx);
  if (np <= 0) return 0;
  if (!pos || !dir || !sigma_H || !sigma_He_corr || !optical_depth) CMIB_FAIL("null argument");
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_pos, *d_dir, *d_sh, *d_she, *d_tau;
  CUDA_OK(sc.in(&d_pos, pos, (size_t)np * 3, s));
  CUDA_OK(sc.in(&d_dir, dir, (size_t)np * 3, s));
  CUDA_OK(sc.in(&d_sh, sigma_H, (size_t)np, s));
  CUDA_OK(sc.in(&d_she, sigma_He_corr, (size_t)np, s));
  CUDA_OK(sc.out(&d_tau, (size_t)np));
  integrate_optical_depth_kernel<<<blocks_for(np, 128), 128, 0, s>>>(ctx->geom, ctx->cells.p, np, d_pos, d_dir, d_sh, d_she,
                                                                    d_tau);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(optical_depth, d_tau, (size_t)np, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_sample_packets(cmib_context *ctx, int64_t n, uint64_t offset, uint64_t seed,
                        uint32_t iteration, double *pos, double *dir, double *nu, double *sigma,
                        double *sigma_He_corr, double *tau) {
  CHECK_CTX(ctx);
  if (ctx->src.n_sources <= 0 && ctx->src.continuous_kind == CONTINUOUS_NONE) CMIB_FAIL("no photon sources set");
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  SamplePacketsParams P;
  P.src = ctx->src;
  P.geom = ctx->geom;
  P.n = n; P.offset = offset; P.seed = seed; P.iteration = iteration;
  CUDA_OK(sc.out(&P.pos, (size_t)n * 3));
  CUDA_OK(sc.out(&P.dir, (size_t)n * 3));
  CUDA_OK(sc.out(&P.nu, (size_t)n));
  CUDA_OK(sc.out(&P.sigma, (size_t)n * NUM_IONS));
  CUDA_OK(sc.out(&P.sigma_He_corr, (size_t)n));
  CUDA_OK(sc.out(&P.tau, (size_t)n));
  sample_packets_kernel<<<blocks_for(n, 128), 128, 0, s>>>(P);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(pos, P.pos, (size_t)n * 3, s));
  CUDA_OK(sc.back(dir, P.dir, (size_t)n * 3, s));
  CUDA_OK(sc.back(nu, P.nu, (size_t)n, s));
  CUDA_OK(sc.back(sigma, P.sigma, (size_t)n * NUM_IONS, s));
  CUDA_OK(sc.back(sigma_He_corr, P.sigma_He_corr, (size_t)n, s));
  CUDA_OK(sc.back(tau, P.tau, (size_t)n, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_cross_sections(cmib_context *ctx, int64_t n, const double *nu, double *sigma) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_nu, *d_sigma;
  CUDA_OK(sc.in(&d_nu, nu, (size_t)n, s));
  CUDA_OK(sc.out(&d_sigma, (size_t)n * NUM_IONS));
  eval_cross_sections_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, ctx->src, d_nu, d_sigma);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(sigma, d_sigma, (size_t)n * NUM_IONS, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_recombination_rates(cmib_context *ctx, int64_t n, const double *T, double *alpha) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_T, *d_a;
  CUDA_OK(sc.in(&d_T, T, (size_t)n, s));
  CUDA_OK(sc.out(&d_a, (size_t)n * NUM_IONS));
  eval_recombination_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, ctx->rr, d_T, d_a);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(alpha, d_a, (size_t)n * NUM_IONS, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_charge_transfer(cmib_context *ctx, int64_t n, const double *T4, double *out) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_T, *d_o;
  CUDA_OK(sc.in(&d_T, T4, (size_t)n, s));
  CUDA_OK(sc.out(&d_o, (size_t)n * 3 * NUM_IONS));
  eval_charge_transfer_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, d_T, d_o);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(out, d_o, (size_t)n * 3 * NUM_IONS, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_line_cooling(cmib_context *ctx, int64_t n, const double *T, const double *ne,
                           const double *abund, double *cooling) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_T, *d_ne, *d_ab, *d_c;
  CUDA_OK(sc.in(&d_T, T, (size_t)n, s));
  CUDA_OK(sc.in(&d_ne, ne, (size_t)n, s));
  CUDA_OK(sc.in(&d_ab, abund, (size_t)n * LC_NUM, s));
  CUDA_OK(sc.out(&d_c, (size_t)n));
  eval_line_cooling_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, d_T, d_ne, d_ab, d_c);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(cooling, d_c, (size_t)n, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_solve5(cmib_context *ctx, int64_t n, double *A, double *B, int32_t *status) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_A, *d_B;
  int32_t *d_s;
  CUDA_OK(sc.in(&d_A, (const double *)A, (size_t)n * 25, s));
  CUDA_OK(sc.in(&d_B, (const double *)B, (size_t)n * 5, s));
  CUDA_OK(sc.out(&d_s, (size_t)n));
  eval_solve5_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, d_A, d_B, d_s);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(A, d_A, (size_t)n * 25, s));
  CUDA_OK(sc.back(B, d_B, (size_t)n * 5, s));
  CUDA_OK(sc.back(status, d_s, (size_t)n, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_reemission_probabilities(cmib_context *ctx, int64_t n, const double *T, double *out) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_T, *d_o;
  CUDA_OK(sc.in(&d_T, T, (size_t)n, s));
  CUDA_OK(sc.out(&d_o, (size_t)n * NUM_REEMIT));
  eval_reemission_probabilities_kernel<<<blocks_for(n, 128), 128, 0, s>>>(n, d_T, d_o);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(out, d_o, (size_t)n * NUM_REEMIT, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

static int eval_state_common(cmib_context *ctx, int solve_T, int64_t n, double jfac, double hfac,
                             const double *J, const double *heat, const double *ndens,
                             const double *T, const double *cr_factor, const double *midz,
                             double *T_out, double *x, double *heat_out) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  EvalStateParams P;
  P.n = n; P.jfac = jfac; P.hfac = hfac;
  for (int k = 0; k < NUM_ELEMENTS; ++k) P.abund[k] = ctx->abund[k];
  P.rr = ctx->rr;
  P.tp = ctx->tp;
  P.solve_temperature = solve_T;
  double *dJ, *dh, *dn, *dT, *dcr, *dmz;
  CUDA_OK(sc.in(&dJ, J, (size_t)n * NUM_IONS, s));
  CUDA_OK(sc.in(&dh, heat, (size_t)n * 2, s));
  CUDA_OK(sc.in(&dn, ndens, (size_t)n, s));
  CUDA_OK(sc.in(&dT, T, (size_t)n, s));
  CUDA_OK(sc.in(&dcr, cr_factor, (size_t)n, s));
  CUDA_OK(sc.in(&dmz, midz, (size_t)n, s));
  P.J = dJ; P.heat = dh; P.ndens = dn; P.T = dT; P.cr_factor = dcr; P.midz = dmz;
  CUDA_OK(sc.out(&P.T_out, (size_t)n));
  CUDA_OK(sc.out(&P.x, (size_t)n * NUM_IONS));
  CUDA_OK(sc.out(&P.heat_out, (size_t)n * 2));
  eval_state_kernel<<<blocks_for(n, 128), 128, 0, s>>>(P);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(T_out, P.T_out, (size_t)n, s));
  CUDA_OK(sc.back(x, P.x, (size_t)n * NUM_IONS, s));
  CUDA_OK(sc.back(heat_out, P.heat_out, (size_t)n * 2, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_eval_ionization_state(cmib_context *ctx, int64_t n, double jfac, double hfac,
                               const double *J, const double *heat, const double *ndens,
                               const double *T, double *x, double *heat_out) {
  return eval_state_common(ctx, 0, n, jfac, hfac, J, heat, ndens, T, nullptr, nullptr, nullptr, x,
                           heat_out);
}

int cmib_eval_temperature(cmib_context *ctx, int64_t n, double jfac, double hfac, const double *J,
                          const double *heat, const double *ndens, const double *T,
                          const double *cr_factor, const double *midz, double *T_out, double *x,
                          double *heat_out) {
  return eval_state_common(ctx, 1, n, jfac, hfac, J, heat, ndens, T, cr_factor, midz, T_out, x,
                           heat_out);
}

int cmib_eval_cooling_heating_balance(cmib_context *ctx, int64_t n, const double *T,
                                      const double *ndens, const double *j, const double *h,
                                      const double *midz, double *h0, double *he0, double *gain,
                                      double *loss, double *metals) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  Scratch sc;
  cudaStream_t s = ctx->stream;
  EvalBalanceParams P;
  P.n = n;
  for (int k = 0; k < NUM_ELEMENTS; ++k) P.abund[k] = ctx->abund[k];
  P.rr = ctx->rr;
  P.pahfac = ctx->tp.pahfac;
  P.crfac = ctx->tp.crfac;
  P.crscale = ctx->tp.crscale;
  double *dT, *dn, *dj, *dh, *dmz;
  CUDA_OK(sc.in(&dT, T, (size_t)n, s));
  CUDA_OK(sc.in(&dn, ndens, (size_t)n, s));
  CUDA_OK(sc.in(&dj, j, (size_t)n * NUM_IONS, s));
  CUDA_OK(sc.in(&dh, h, (size_t)n * 2, s));
  CUDA_OK(sc.in(&dmz, midz, (size_t)n, s));
  P.T = dT; P.ndens = dn; P.j = dj; P.h = dh; P.midz = dmz;
  CUDA_OK(sc.out(&P.h0, (size_t)n));
  CUDA_OK(sc.out(&P.he0, (size_t)n));
  CUDA_OK(sc.out(&P.gain, (size_t)n));
  CUDA_OK(sc.out(&P.loss, (size_t)n));
  CUDA_OK(sc.out(&P.metals, (size_t)n * 12));
  eval_balance_kernel<<<blocks_for(n, 128), 128, 0, s>>>(P);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(h0, P.h0, (size_t)n, s));
  CUDA_OK(sc.back(he0, P.he0, (size_t)n, s));
  CUDA_OK(sc.back(gain, P.gain, (size_t)n, s));
  CUDA_OK(sc.back(loss, P.loss, (size_t)n, s));
  CUDA_OK(sc.back(metals, P.metals, (size_t)n * 12, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

int cmib_get_spectrum_tables(cmib_context *ctx, int which, double *a, double *b, double *c) {
  if (!ctx) CMIB_FAIL("null context");
  auto copy = [](double *dst, const std::vector<double> &v) {
    if (dst && !v.empty()) memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  switch (which) {
  case 0:
    if (ctx->h_planck.empty()) CMIB_FAIL("no Planck spectrum set");
    copy(a, ctx->h_planck);
    break;
  case 1:
    if (ctx->h_hlyc_cdf.empty()) CMIB_FAIL("no Physical reemission handler set");
    copy(a, ctx->h_hlyc_freq); copy(b, ctx->h_hlyc_temp); copy(c, ctx->h_hlyc_cdf);
    break;
  case 2:
    if (ctx->h_helyc_cdf.empty()) CMIB_FAIL("no Physical reemission handler set");
    copy(a, ctx->h_helyc_freq); copy(b, ctx->h_helyc_temp); copy(c, ctx->h_helyc_cdf);
    break;
  case 3:
    if (ctx->h_he2pc_cdf.empty()) CMIB_FAIL("no Physical reemission handler set");
    copy(a, ctx->h_he2pc_freq); copy(b, ctx->h_he2pc_cdf);
    break;
  default:
    CMIB_FAIL("unknown table %d", which);
  }
  return 0;
}

int cmib_sample_spectrum(cmib_context *ctx, int which, double temperature, uint64_t seed, int64_t n,
                         double *nu) {
  CHECK_CTX(ctx);
  if (n <= 0) return 0;
  if (which != 0 && which != 4 && ctx->src.reemission_kind != REEMISSION_PHYSICAL)
    CMIB_FAIL("diffuse spectra need the Physical reemission handler");
  Scratch sc;
  cudaStream_t s = ctx->stream;
  double *d_nu;
  CUDA_OK(sc.out(&d_nu, (size_t)n));
  sample_spectrum_kernel<<<blocks_for(n, 128), 128, 0, s>>>(ctx->src, which, temperature, seed, n, d_nu);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  CUDA_OK(sc.back(nu, d_nu, (size_t)n, s));
  CUDA_OK(cudaStreamSynchronize(s));
  return 0;
}

} /* extern "C" */
