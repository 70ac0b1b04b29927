/*
 * cmib_api.cu — implementation of the C ABI declared in include/cmib.h.
 *
 * Host side is thin: it owns device buffers, builds the tabulated spectra on the
 * host (spectrum_tables.hpp), and launches the kernels of kernels.cuh on the
 * context's stream.  There is no CPU compute path: every entry point that does
 * work requires a CUDA device and fails loudly otherwise.
 */
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <deque>
#include <vector>

#include <dlfcn.h>
#include <nccl.h> /* types only: libnccl.so.2 is bound at run time (NcclApi), nothing is linked */


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

/* NCCL entry points, resolved on first use.  A python process that imported torch has torch's bundled
 * libnccl.so.2 loaded already and dlopen returns that one; otherwise the system library. */
struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  std::string error;
  static NcclApi &get() {
    static NcclApi api = load();
    return api;
  }

private:
  static NcclApi load() {
    NcclApi a;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      a.error = std::string("multi-GPU runs need NCCL: ") + dlerror();
      return a;
    }
#define CMIB_NCCL_SYM(name) a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name))
    CMIB_NCCL_SYM(GetUniqueId); CMIB_NCCL_SYM(CommInitRank); CMIB_NCCL_SYM(CommInitAll); CMIB_NCCL_SYM(CommDestroy);
    CMIB_NCCL_SYM(AllReduce); CMIB_NCCL_SYM(AllGather); CMIB_NCCL_SYM(Reduce); CMIB_NCCL_SYM(Broadcast); CMIB_NCCL_SYM(GroupStart);
    CMIB_NCCL_SYM(GroupEnd); CMIB_NCCL_SYM(GetErrorString);
#undef CMIB_NCCL_SYM
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommInitAll || !a.CommDestroy || !a.AllReduce || !a.AllGather || !a.Reduce ||
        !a.Broadcast || !a.GroupStart || !a.GroupEnd || !a.GetErrorString)
      a.error = "libnccl lacks the expected entry points";
    return a;
  }
};
#define NCCL_OK(expr)                                                                                      \
  do {                                                                                                     \
    ncclResult_t r_ = (expr);                                                                              \
    if (r_ != ncclSuccess) CMIB_FAIL("%s failed: %s", #expr, NcclApi::get().GetErrorString(r_));          \
  } while (0)

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
  /* A shoot runs as one or two LANES: each lane owns a set of queues, a control block and a stream and walks its
   * share of the packets through rounds of decide -> prepare -> (sort) -> march.  With two lanes the kernels are
   * launched on half-size grids and the lanes run out of phase, so that the emission of one lane (FP64 / issue
   * bound) executes on the same SMs as the walk of the other (L1TEX / latency bound) instead of before it. */
  struct Lane {
    uint64_t queue_capacity = 0;
    int queue_mode = -1;
    DevBuf<double> mq, rq, eq;
    DevBuf<unsigned long long> ctl;
    DevBuf<uint32_t> sort_key, sort_rank, sort_order, sort_hist, sort_offs, sort_block_sums; /* coherent march: counting sort */
    unsigned long long *h_ctl = nullptr; /* pinned mirrors of the control block: one per group of rounds in flight */
    cudaStream_t stream = nullptr;       /* lane 0: the context's stream */
    cudaEvent_t group_done[2] = {nullptr, nullptr};
    std::vector<cudaEvent_t> ev_pool;    /* optional per-kernel timing */
    size_t ev_used = 0;
  };
  Lane lane[2];
  cudaEvent_t fork_ev = nullptr, origin_ev = nullptr;
  int shoot_lanes = 1;        /* lanes of the last shoot */
  double overlap_ms = 0.;     /* time of the last shoot during which an emission kernel ran beside a march kernel */
  DevBuf<uint32_t> d_src_cell;  /* packed cell indices of the sources */
  std::vector<uint32_t> h_src_cell;
  int hot_replicas = 0;
  DevBuf<unsigned long long> upd_counter; /* next unprocessed cell of update_temperature_kernel */
  int update_blocks_per_sm[2] = {0, 0};
  /* march-queue order: 0 emission order, 2 coherent march = queue read in key order + in-warp sums,
   * -1 (default) by rule: 2 when cells + accumulators do not fit in L2 and the shoot is large, else 0.
   * (Round 1 timed the two orders against each other inside live shoots because its coherent H-only kernel
   * lost on single-source grids; march_lean_kernel wins on every HBM-resident grid measured, so the choice
   * is a deterministic function of the working set: step times are reproducible.)  Both orders shoot the
   * same packets; only the order of the atomic adds differs. */
  int sort_mode = -1;
  /* multi-GPU (MPICommunicator of the reference): one communicator rank per context */
  ncclComm_t comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  cudaEvent_t xev[4] = {nullptr, nullptr, nullptr, nullptr};
  DevBuf<double> xchg_pack, xchg_all; /* packs of the owned chunks: own, and of all ranks (gathers) */
  double walk_cells = 0.; /* mean walk length (cells) of the last large shoot: sizes the direction bins of the sort key */
  size_t l2_bytes = 0;
  int march_blocks_per_sm[2][2] = {{0, 0}, {0, 0}}; /* [layout][plain, coherent] */
  int lean_grid[2][2][2][2] = {};                   /* resident CTAs per SM of the march_lean_kernel variants */
  int prep_blocks_per_sm[2] = {0, 0};
  int decide_blocks_per_sm = 0;
  uint64_t shoot_rounds = 0;
  /* optional per-kernel timing of the shoot (CUDA events on the context's stream) */
  bool timing = false;
  double prepare_ms = 0., march_ms = 0.;
  double nu_H = 0., nu_He = 0.;
  int conventions = 0; /* CMIB_CONVENTIONS_* */
  /* H-only planes: has anything been shot since the last reset that can add a heating term?  A monochromatic source at
   * exactly nu_H without re-emission adds none (nu - nu_H == 0), and then the exchange of a multi-GPU iteration sums
   * the J_H plane only (half the bytes: 134 instead of 268 MB on a 256^3 grid) */
  bool heat_possibly_written = true;
  /* the abundance factors the packets carry (SourceModel::fold) follow the abundances and the conventions */
  void apply_conventions() {
    static const int element_of_ion[NUM_IONS] = {-1, EL_He, EL_C, EL_C, EL_N, EL_N, EL_N, EL_O, EL_O, EL_Ne, EL_Ne, EL_S, EL_S, EL_S};
    for (int ion = 0; ion < NUM_IONS; ++ion)
      src.fold[ion] = (conventions == 1 && ion > 0) ? abund[element_of_ion[ion]] : 1.;
    src.A_He = abund[EL_He];
    src.A_He_reemit = (conventions == 1) ? 1. : abund[EL_He];
  }

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
  /* the accumulator buffer: counters + per-cell accumulators (acc_main_doubles: what collectives move and callers
   * see), then the hot-cell replicas of the shoot (hot_doubles; always zero outside a shoot).  One allocation, so
   * that the walk forms every accumulator address from one base and a 32-bit index (march_coherent.cuh). */
  size_t acc_main_doubles(int mode) const {
    size_t n = (mode != ACC_HONLY) ? ACC_COUNTERS + (size_t)geom.ncells * 16
                                   : ACC_COUNTERS + 64 + (size_t)geom.ncells * (honly_planar() ? 2 : (size_t)honly_cell_stride());
    return (n + 15) & ~(size_t)15;
  }
  size_t hot_doubles() const { return hot_replicas > 0 ? (size_t)hot_replicas * HOT_MAX_SOURCES * HOT_CELLS * HOT_STRIDE : 0; }
  size_t acc_doubles(int mode) const { return acc_main_doubles(mode) + hot_doubles(); }
  double *hot_acc() const { return hot_replicas > 0 ? acc.p + acc_main_doubles(acc_mode) : nullptr; }
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
    ctx->heat_possibly_written = false;
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

/* packets per round of the wavefront pipeline.  Emission order: 64 Mi (21 GB of queues in the full layout, of 180 GB):
 * every round ends in a tail in which the last long walks keep the whole machine waiting, so fewer, larger rounds win
 * — lexingtonHII20, 1e8 packets: 4 Mi 124 ms (r01), 16 Mi 96.9 ms, 32 Mi 95.7, 64 Mi 95.2, 128 Mi 95.1.  Coherent
 * march: what the sort key needs (shoot_wavefront). */
uint64_t default_queue_capacity(bool coherent, uint64_t coherent_round) {
  const char *e = getenv("CMIB_QUEUE_CAPACITY");
  if (e) {
    const long long v = atoll(e);
    if (v >= 1024) return (uint64_t)v;
  }
  return coherent ? coherent_round : (1ull << 26);
}

/* measure of a union of intervals (sorted in place), and of the intersection of two such unions */
void merge_intervals(std::vector<std::pair<double, double>> &v) {
  std::sort(v.begin(), v.end());
  size_t n = 0;
  for (size_t i = 0; i < v.size(); ++i) {
    if (n > 0 && v[i].first <= v[n - 1].second) v[n - 1].second = std::max(v[n - 1].second, v[i].second);
    else v[n++] = v[i];
  }
  v.resize(n);
}
double interval_measure(const std::vector<std::pair<double, double>> &v) {
  double m = 0.;
  for (const auto &i : v) m += i.second - i.first;
  return m;
}
double interval_overlap(const std::vector<std::pair<double, double>> &a, const std::vector<std::pair<double, double>> &b) {
  double m = 0.;
  size_t i = 0, j = 0;
  while (i < a.size() && j < b.size()) {
    const double lo = std::max(a[i].first, b[j].first), hi = std::min(a[i].second, b[j].second);
    if (hi > lo) m += hi - lo;
    if (a[i].second < b[j].second) ++i; else ++j;
  }
  return m;
}

/* one cmib_shoot call on the wavefront path: rounds of prepare -> march until the queues run dry, on one or two
 * lanes (cmib_context::Lane).  The host reads a lane's control block back once per group of rounds. */
int shoot_wavefront(cmib_context *ctx, const ShootParams &P) {
  const int mode = ctx->acc_mode;
  /* order of the march queue (cmib_context::sort_mode; CMIB_SORT overrides for A/B runs) */
  int sort = ctx->sort_mode;
  if (const char *e = getenv("CMIB_SORT")) sort = atoi(e);
  if (sort < 0) {
    const size_t working_set = (size_t)ctx->geom.ncells * ((mode == ACC_HONLY ? 16 : sizeof(CellOpacity)) + (mode == ACC_HONLY ? 16 : 128));
    sort = (working_set > ctx->l2_bytes && P.n_packets >= (1ull << 20)) ? 2 : 0;
  }
  /* coherent march: the key wants source + direction bits (below) + a few optical-depth bits, and ~4 packets per
   * bin: that sets the size of a round, between 16 Mi (one source) and 64 Mi packets (16 sources: clumpy 256^3
   * 7.7e8 -> 9.1e8 packets/s from 16 Mi to 64 Mi, profiles/r02_lean_march.md) */
  int src_bits = 0, dir_bits = 0;
  uint64_t coherent_round = 1ull << 26;
  if (sort == 2 || sort == 1) {
    while ((1ll << src_bits) < (long long)P.src.n_sources) ++src_bits;
    /* how many direction bins?  Packets of one bin should still cross the same cells at the END of their walk:
     * bin width x walk length ~ one cell, i.e. 4 pi L^2 bins per source for walks of L cells.  L comes from the
     * previous large shoot of this context (cell crossings per walk / 1.5: an isotropic direction crosses
     * |dx| + |dy| + |dz| = 1.5 walls per cell of path on average), else half the grid.  Measured (B200, 1.6e7
     * packets per round): stromgren 256^3, L = 85: 16 direction + 6 depth bits 11.0 ms, 18 + 4 11.5, 22 + 0 13.1;
     * clumpy 256^3, 16 sources, L = 106: 18 + 0 19.7 ms, 16 + 2 20.2, 14 + 4 22.2 (profiles/r02_lean_march.md). */
    double walk = 0.5 * (double)std::max(P.geom.ncell[0], std::max(P.geom.ncell[1], P.geom.ncell[2]));
    if (ctx->walk_cells > 0.) walk = ctx->walk_cells;
    dir_bits = 2 * (int)std::lround(0.5 * std::log2(4. * 3.14159265358979 * walk * walk)); /* nearest even */
    if (const char *e = getenv("CMIB_DIR_BITS")) dir_bits = atoi(e) & ~1;
    if (dir_bits > 22) dir_bits = 22;
    if (dir_bits < 8) dir_bits = 8;
    if (src_bits + dir_bits > SORT_MAX_KEY_BITS) dir_bits = (SORT_MAX_KEY_BITS - src_bits) & ~1;
    const int want_bits = std::min(src_bits + dir_bits + 4, 24) + 2;
    coherent_round = std::min<uint64_t>(std::max<uint64_t>(1ull << want_bits, 1ull << 24), 1ull << 26);
  }
  /* lanes: one.  Two lanes (CMIB_LANES=2: half grids, out of phase, so that emission and sort of one lane run beside
   * the walk of the other) were worth +8 % on the one-source 256^3 grid while the queue writes of prepare_kernel still
   * pushed the cone's cells out of L2 (stromgren 256^3, 1e9 packets: 764 -> 704 ms); with streaming queue stores the
   * serial emission costs 5 % of the shoot and two lanes return 1 % (696 vs 689 ms), nothing on clumpy 256^3 (latency
   * bound, needs the largest rounds) and nothing on lexingtonHII20 64^3 (walk and emission both live on L1TEX:
   * 98.4 vs 98.9 ms) — profiles/r02_lanes.md.  One lane keeps the march kernel alone on the machine, which is also
   * what its roofline figure assumes. */
  int nlanes = 1;
  if (const char *e = getenv("CMIB_LANES")) nlanes = atoi(e) == 2 ? 2 : 1;
  if (P.n_packets < 2048) nlanes = 1;
  uint64_t lane_n[2] = {P.n_packets, 0};
  if (nlanes == 2) {
    lane_n[1] = P.n_packets / 2;
    lane_n[0] = P.n_packets - lane_n[1];
  }
  const int nf = (mode == ACC_HONLY) ? MarchQueueLayout<ACC_HONLY>::NFIELDS : MarchQueueLayout<ACC_FULL>::NFIELDS;
  uint64_t lane_cap[2] = {0, 0};
  for (int l = 0; l < nlanes; ++l) {
    cmib_context::Lane &L = ctx->lane[l];
    if (!L.stream) {
      if (l == 0) L.stream = ctx->stream;
      else CUDA_OK(cudaStreamCreateWithFlags(&L.stream, cudaStreamNonBlocking));
    }
    uint64_t cap = default_queue_capacity(sort != 0, coherent_round);
    if (lane_n[l] < cap) cap = (lane_n[l] + 1023) / 1024 * 1024;
    if (L.queue_capacity < cap || L.queue_mode != mode) {
      CUDA_OK(L.mq.resize((size_t)nf * cap));
      CUDA_OK(L.rq.resize((size_t)RQ_NFIELDS * cap));
      CUDA_OK(L.eq.resize((size_t)EQ_NFIELDS * cap));
      L.queue_capacity = cap;
      L.queue_mode = mode;
    }
    lane_cap[l] = L.queue_capacity;
    if (!L.ctl.p) {
      CUDA_OK(L.ctl.resize(CTL_WORDS));
      CUDA_OK(cudaMemsetAsync(L.ctl.p, 0, CTL_WORDS * sizeof(unsigned long long), L.stream));
      CUDA_OK(cudaMallocHost((void **)&L.h_ctl, 2 * CTL_WORDS * sizeof(unsigned long long)));
      for (cudaEvent_t &e : L.group_done) CUDA_OK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
  }
  if (!ctx->fork_ev) {
    CUDA_OK(cudaEventCreateWithFlags(&ctx->fork_ev, cudaEventDisableTiming));
    CUDA_OK(cudaEventCreate(&ctx->origin_ev));
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
  cudaStream_t s0 = ctx->stream;
  WavefrontParams W;
  W.sp = P;
  W.acc_j = P.acc + ACC_COUNTERS + P.honly_offset;
  W.hot_index0 = P.hot_acc ? (uint32_t)(P.hot_acc - W.acc_j) : 0xffffffffu;
  for (int d = 0; d < 3; ++d) W.lean_n16[d] = 16u * (uint32_t)P.geom.ncell[d];
  W.lean_k[0] = P.geom.ncell[1] * P.geom.ncell[2];
  W.lean_k[1] = P.geom.ncell[2];
  W.agg = 1;
  if (sort == 1) { sort = 2; W.agg = 0; } /* A/B: ordered queue, plain kernel */
  W.sort = sort;
  W.key = W.rank = W.order = W.hist = W.offs = W.block_sums = nullptr;
  W.nbins = 0; W.fine_dir_bits = 22; W.fine_key_bits = 22; W.tau_bits = 0; W.chunk_stride = 1;
  /* entries a round of lane l may hold: the lane's capacity; CMIB_LANE_ROUNDS = r asks for at least r rounds per
   * lane (A/B runs); with two lanes the first round of lane 1 is half a round, which puts the lanes out of phase:
   * the emission of one lane then runs beside the walk of the other */
  uint64_t lane_fill[2] = {lane_cap[0], lane_cap[1]};
  {
    int min_rounds = 1;
    if (const char *e = getenv("CMIB_LANE_ROUNDS")) min_rounds = atoi(e) > 1 ? atoi(e) : 1;
    for (int l = 0; l < nlanes; ++l) {
      const uint64_t per = ((lane_n[l] + min_rounds - 1) / min_rounds + 1023) / 1024 * 1024;
      if (per < lane_fill[l]) lane_fill[l] = per;
      if (lane_fill[l] < 1024) lane_fill[l] = 1024;
    }
  }
  auto round_fill = [&](int l, uint64_t round) -> uint64_t {
    if (nlanes == 2 && l == 1 && round == 0) {
      const uint64_t first = std::min(lane_fill[1], lane_n[1]);
      return std::max<uint64_t>(1024, (first / 2 + 1023) / 1024 * 1024);
    }
    return lane_fill[l];
  };
  bool sort_reemitted_rounds = false;
  if (sort == 2) {
    /* key = source | direction | optical depth (wavefront.cuh): 0 | key for primaries, 1 | position key for
     * re-emitted packets.  The bits that are left (of ~log2(packets per round / 4)) order the packets of a direction
     * bin by sampled optical depth, so that the 8 lanes of a refill group are absorbed close together. */
    int budget = 2; /* ~4 packets per bin */
    const uint64_t per_round = std::min(lane_n[0], lane_fill[0]);
    while ((1ull << budget) < per_round) ++budget;
    budget -= 2;
    if (budget > SORT_MAX_KEY_BITS) budget = SORT_MAX_KEY_BITS;
    if (const char *e = getenv("CMIB_KEY_BITS")) budget = atoi(e) > SORT_MAX_KEY_BITS ? SORT_MAX_KEY_BITS : atoi(e);
    int tau_bits = budget - src_bits - dir_bits;
    if (const char *e = getenv("CMIB_TAU_BITS")) tau_bits = atoi(e);
    if (tau_bits < 0) tau_bits = 0;
    if (tau_bits > 8) tau_bits = 8;
    if (src_bits + dir_bits + tau_bits > SORT_MAX_KEY_BITS) tau_bits = SORT_MAX_KEY_BITS - src_bits - dir_bits;
    if (dir_bits < 2) W.sort = sort = 0; /* no room for a direction next to that many sources */
    W.fine_dir_bits = dir_bits;
    W.tau_bits = tau_bits;
    W.fine_key_bits = dir_bits + src_bits + tau_bits;
    if (W.fine_key_bits < 12) W.fine_key_bits = 12; /* nbins a multiple of the scan tile */
  }
  if (sort == 2) {
    if (const char *e = getenv("CMIB_CHUNK_STRIDE")) W.chunk_stride = (uint32_t)atoll(e); /* e.g. the prime 1000003 */
    if (W.chunk_stride < 1u || std::max(lane_cap[0], lane_cap[1]) / MARCH_CHUNK >= W.chunk_stride) W.chunk_stride = 1u;
    if (const char *e = getenv("CMIB_AGG")) W.agg = atoi(e) != 0;
    if (const char *e = getenv("CMIB_SORT_REEMITTED")) sort_reemitted_rounds = atoi(e) != 0;
    W.nbins = 2u << W.fine_key_bits;
    for (int l = 0; l < nlanes; ++l) {
      cmib_context::Lane &L = ctx->lane[l];
      if (L.sort_key.n < lane_cap[l]) {
        CUDA_OK(L.sort_key.resize(lane_cap[l]));
        CUDA_OK(L.sort_rank.resize(lane_cap[l]));
        CUDA_OK(L.sort_order.resize(lane_cap[l]));
      }
      if (L.sort_hist.n < W.nbins) {
        CUDA_OK(L.sort_hist.resize(W.nbins));
        CUDA_OK(L.sort_offs.resize(W.nbins));
        CUDA_OK(L.sort_block_sums.resize(SORT_MAX_TILES));
        CUDA_OK(cudaMemsetAsync(L.sort_hist.p, 0, W.nbins * sizeof(uint32_t), s0));
      }
    }
  }
  /* persistent grids: exactly the CTAs that are resident at once (a partial second wave of a
   * grid-stride kernel runs at a fraction of the machine); with two lanes each lane launches half of them */
  if (ctx->prep_blocks_per_sm[0] == 0) {
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prep_blocks_per_sm[ACC_FULL], prepare_kernel<ACC_FULL>, 256, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->prep_blocks_per_sm[ACC_HONLY], prepare_kernel<ACC_HONLY>, 256, 0));
    CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctx->decide_blocks_per_sm, reemit_decide_kernel, 256, 0));
  }
  auto lane_grid = [&](unsigned full) -> unsigned { return nlanes == 2 ? (full + 1u) / 2u : full; };
  const unsigned prep_grid = lane_grid((unsigned)(ctx->sm_count * (ctx->prep_blocks_per_sm[mode] > 0 ? ctx->prep_blocks_per_sm[mode] : 1)));
  const unsigned decide_grid = lane_grid((unsigned)(ctx->sm_count * (ctx->decide_blocks_per_sm > 0 ? ctx->decide_blocks_per_sm : 1)));
  unsigned march_grids[2];
  for (int a = 0; a < 2; ++a) {
    int bpm = ctx->march_blocks_per_sm[mode][a];
    if (const char *e = getenv("CMIB_MARCH_BLOCKS_PER_SM")) bpm = atoi(e);
    if (bpm < 1) bpm = 1;
    march_grids[a] = lane_grid((unsigned)(ctx->sm_count * bpm));
  }
  const int sort_cfg = W.sort;
  /* H-only coherent walk: march_lean_kernel (CMIB_LEAN=0: the r01 kernel, for A/B runs) */
  int lean_cfg = 1, lean_steps = 3;
  if (const char *e = getenv("CMIB_LEAN")) lean_cfg = atoi(e) != 0;
  if (const char *e = getenv("CMIB_LEAN_STEPS")) lean_steps = atoi(e) == 2 ? 2 : 3;
  const size_t lean_smem = lean_smem_bytes(P.geom);
  if (lean_smem > 200 * 1024) lean_cfg = 0; /* wall tables of > ~11000 cells per axis sum: the r01 kernel */
  /* march_lean_kernel addresses accumulators (and the hot-cell replicas behind them) by a 32-bit index in doubles */
  if (ctx->acc_doubles(mode) >= (size_t)0xffffffffu) lean_cfg = 0;
  /* a heat term exists unless every packet of the shoot sits exactly at the threshold: a monochromatic
   * source at nu_H without re-emission (then nu - nu_H == 0 and the r01 kernel skipped the add at run time) */
  const int lean_heat = !(P.src.spectrum.kind == SPECTRUM_MONOCHROMATIC && P.src.spectrum.mono_frequency == P.nu_H &&
                          P.src.reemission_kind == REEMISSION_NONE && P.src.continuous_kind == CONTINUOUS_NONE);
  const int lean_periodic = (P.geom.periodic[0] | P.geom.periodic[1] | P.geom.periodic[2]) ? 1 : 0;
  /* ... and then every packet carries the same sigma_H (FixedValue / Bimodal tables, source.cuh packet_cross_sections:
   * the H-only layout excludes Verner) and the weight of the discrete sources */
  W.uni_sigH = (P.src.xs_kind == XS_BIMODAL && !(P.src.spectrum.mono_frequency < P.src.xs_limit)) ? P.src.xs_high[0] : P.src.xs_fixed[0];
  W.uni_w = P.src.discrete_weight;
  /* march_kernel<.., PRE>: request the next cell record one pass ahead.  Measured (B200): pays in the
   * coherent kernel, whose in-warp sums sit between request and use (clumpy 256^3 30.1 -> 26.3 ms);
   * the plain kernel is bound by L1TEX lanes, not latency (no gain; -16 % with the full layout's spills) */
  int prefetch_cfg = -1;
  if (const char *e = getenv("CMIB_PREFETCH")) prefetch_cfg = atoi(e) != 0;
  const bool can_reemit = (P.src.reemission_kind != REEMISSION_NONE);
  int tail_cfg = 1;
  if (const char *e = getenv("CMIB_TAIL")) tail_cfg = atoi(e) != 0;

  /* ---- per-lane state of this shoot ---- */
  struct Run {
    WavefrontParams W;
    uint64_t rounds = 0;      /* rounds enqueued */
    int enq = 0, fin = 0;     /* groups enqueued / read back */
    uint64_t group_first[2] = {0, 0};
    int group_rounds[2] = {0, 0};
    bool done = false, primaries_left = true;
  } run[2];
  if (ctx->timing) CUDA_OK(cudaEventRecord(ctx->origin_ev, s0));
  shoot_begin_kernel<<<1, 1, 0, s0>>>(ctx->lane[0].ctl.p, P.acc);
  ++g_launches;
  if (nlanes == 2) {
    /* lane 1 starts after everything that is queued on the context's stream */
    CUDA_OK(cudaEventRecord(ctx->fork_ev, s0));
    CUDA_OK(cudaStreamWaitEvent(ctx->lane[1].stream, ctx->fork_ev, 0));
  }
  for (int l = 0; l < nlanes; ++l) {
    cmib_context::Lane &L = ctx->lane[l];
    Run &R = run[l];
    R.W = W;
    R.W.sp.packet_offset = P.packet_offset + (l == 0 ? 0 : lane_n[0]);
    R.W.sp.n_packets = lane_n[l];
    R.W.ctl = L.ctl.p;
    R.W.mq = L.mq.p;
    R.W.rq = L.rq.p;
    R.W.eq = L.eq.p;
    R.W.capacity = lane_cap[l];
    R.W.fill = lane_fill[l];
    if (sort_cfg == 2) {
      R.W.key = L.sort_key.p; R.W.rank = L.sort_rank.p; R.W.order = L.sort_order.p;
      R.W.hist = L.sort_hist.p; R.W.offs = L.sort_offs.p; R.W.block_sums = L.sort_block_sums.p;
    }
    L.ev_used = 0;
    /* control block: everything zero but the primaries to emit (lane 0 keeps what shoot_begin_kernel wrote) */
    memset(L.h_ctl, 0, CTL_WORDS * sizeof(unsigned long long));
    L.h_ctl[CTL_REMAINING] = lane_n[l];
    CUDA_OK(cudaMemcpyAsync(L.ctl.p, L.h_ctl, CTL_CROSSINGS0 * sizeof(unsigned long long), cudaMemcpyHostToDevice, L.stream));
  }
  auto stamp = [&](cmib_context::Lane &L) {
    if (!ctx->timing) return;
    if (L.ev_used == L.ev_pool.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      L.ev_pool.push_back(e);
    }
    cudaEventRecord(L.ev_pool[L.ev_used++], L.stream);
  };
  /* on any failure below nothing may stay in flight, and the hot-cell replicas and bin counts must not leak
   * into the next shoot */
  auto fail_cleanup = [&]() {
    for (int l = 0; l < nlanes; ++l) cudaStreamSynchronize(ctx->lane[l].stream);
    if (P.hot_replicas > 0 && P.hot_acc) cudaMemsetAsync(P.hot_acc, 0, ctx->hot_doubles() * sizeof(double), s0);
    for (int l = 0; l < nlanes; ++l)
      if (run[l].W.hist) cudaMemsetAsync(run[l].W.hist, 0, W.nbins * sizeof(uint32_t), s0); /* bin counts are zero between rounds */
    cudaStreamSynchronize(s0);
  };

  auto enqueue_round = [&](int l) -> int {
    cmib_context::Lane &L = ctx->lane[l];
    Run &R = run[l];
    WavefrontParams &Wl = R.W;
    cudaStream_t s = L.stream;
    const uint64_t round = R.rounds;
    /* rounds without primaries (re-emitted packets start anywhere) run unsorted through the plain kernel */
    if (sort_cfg == 2) Wl.sort = (R.primaries_left || sort_reemitted_rounds) ? 2 : 0;
    const int sort = Wl.sort;
    Wl.fill = round_fill(l, round);
    stamp(L);
    if (can_reemit && round > 0) {
      if (tail_cfg) {
        /* the last generations of re-emitted packets in one launch (a no-op until few enough are left) */
        const unsigned tail_grid = (unsigned)((TAIL_MAX + TAIL_BLOCK - 1) / TAIL_BLOCK);
        if (mode == ACC_HONLY) tail_kernel<ACC_HONLY><<<tail_grid, TAIL_BLOCK, 0, s>>>(Wl);
        else tail_kernel<ACC_FULL><<<tail_grid, TAIL_BLOCK, 0, s>>>(Wl);
        ++g_launches;
      }
      reemit_decide_kernel<<<decide_grid, 256, 0, s>>>(Wl);
      ++g_launches;
    }
    if (mode == ACC_HONLY) prepare_kernel<ACC_HONLY><<<prep_grid, 256, 0, s>>>(Wl);
    else prepare_kernel<ACC_FULL><<<prep_grid, 256, 0, s>>>(Wl);
    stamp(L);
    advance_after_prepare_kernel<<<1, 1, 0, s>>>(Wl.ctl, Wl.fill);
    if (sort == 2) {
      /* counting sort: prepare_kernel counted the bins and handed out tickets */
      const unsigned ntiles = Wl.nbins / SORT_SCAN_TILE;
      sort_scan_tiles_kernel<<<ntiles, SORT_SCAN_BLOCK, 0, s>>>(Wl.hist, Wl.block_sums);
      sort_scan_sums_kernel<<<1, SORT_SCAN_BLOCK, 0, s>>>(Wl.block_sums, ntiles);
      sort_scan_offsets_kernel<<<ntiles, SORT_SCAN_BLOCK, 0, s>>>(Wl.hist, Wl.block_sums, Wl.offs);
      sort_scatter_kernel<<<prep_grid, 256, 0, s>>>(Wl.ctl, Wl.key, Wl.rank, Wl.offs, Wl.order);
      g_launches += 4;
    }
    stamp(L);
    {
      const bool agg = (sort == 2 && Wl.agg);
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
        k<<<lane_grid((unsigned)(ctx->sm_count * bpm)), MARCH_BLOCK, lean_smem, s>>>(Wl);
      } else
#define CMIB_LAUNCH_MARCH(M, A, R) march_kernel<M, A, R><<<march_grid, MARCH_BLOCK, 0, s>>>(Wl)
      if (mode == ACC_HONLY) {
        if (agg) { if (prefetch) CMIB_LAUNCH_MARCH(ACC_HONLY, true, true); else CMIB_LAUNCH_MARCH(ACC_HONLY, true, false); }
        else { if (prefetch) CMIB_LAUNCH_MARCH(ACC_HONLY, false, true); else CMIB_LAUNCH_MARCH(ACC_HONLY, false, false); }
      } else {
        if (agg) { if (prefetch) CMIB_LAUNCH_MARCH(ACC_FULL, true, true); else CMIB_LAUNCH_MARCH(ACC_FULL, true, false); }
        else { if (prefetch) CMIB_LAUNCH_MARCH(ACC_FULL, false, true); else CMIB_LAUNCH_MARCH(ACC_FULL, false, false); }
      }
#undef CMIB_LAUNCH_MARCH
    }
    stamp(L);
    advance_after_march_kernel<<<1, 1, 0, s>>>(Wl.ctl, P.acc);
    g_launches += 4;
    ++R.rounds;
    return 0;
  };
  /* a group: `n` rounds, then the control block comes back into one of the lane's two pinned mirrors */
  auto enqueue_group = [&](int l, int n) -> int {
    cmib_context::Lane &L = ctx->lane[l];
    Run &R = run[l];
    const int slot = R.enq & 1;
    R.group_first[slot] = R.rounds;
    R.group_rounds[slot] = n;
    for (int k = 0; k < n; ++k)
      if (enqueue_round(l)) return 1;
    if (cudaGetLastError() != cudaSuccess) { fail_cleanup(); CMIB_FAIL("a kernel launch of the shoot failed"); }
    CUDA_OK(cudaMemcpyAsync(L.h_ctl + slot * CTL_WORDS, L.ctl.p, CTL_WORDS * sizeof(unsigned long long), cudaMemcpyDeviceToHost, L.stream));
    CUDA_OK(cudaEventRecord(L.group_done[slot], L.stream));
    ++R.enq;
    return 0;
  };
  /* rounds a lane needs to emit its primaries when nothing is re-emitted (re-emitted packets take room: then
   * this is a lower bound); one more round finds the queues empty and reports the end */
  auto rounds_for_primaries = [&](int l) -> int {
    uint64_t rem = lane_n[l], r = 0;
    while (rem > 0 && r < 100000) {
      const uint64_t f = round_fill(l, r);
      rem -= std::min(rem, f);
      ++r;
    }
    return (int)r;
  };
  const int group = 4;
  const uint64_t max_rounds = 100000; /* every round emits min(remaining, room) primaries or shrinks the
                                       * re-emission population; a sane configuration stays far below */
  std::deque<int> inflight; /* lanes in the order their groups were enqueued */
  for (int l = 0; l < nlanes; ++l) {
    if (enqueue_group(l, std::min(rounds_for_primaries(l) + 1, (int)CTL_STATUS_SLOTS - 1))) return 1;
    inflight.push_back(l);
  }
  if (can_reemit) /* the number of rounds is not known: keep one group ahead of the read-back */
    for (int l = 0; l < nlanes; ++l) {
      if (enqueue_group(l, group)) return 1;
      inflight.push_back(l);
    }
  while (!inflight.empty()) {
    const int l = inflight.front();
    inflight.pop_front();
    cmib_context::Lane &L = ctx->lane[l];
    Run &R = run[l];
    const int slot = R.fin & 1;
    CUDA_OK(cudaEventSynchronize(L.group_done[slot]));
    ++R.fin;
    const unsigned long long *h = L.h_ctl + slot * CTL_WORDS;
    if (h[CTL_ERROR]) { fail_cleanup(); CMIB_FAIL("march kernel exceeded its pass limit (internal error)"); }
    for (int k = 0; k < R.group_rounds[slot]; ++k)
      if (h[CTL_STATUS + ((R.group_first[slot] + k) % CTL_STATUS_SLOTS)] == 0) R.done = true;
    if (h[CTL_REMAINING] == 0) R.primaries_left = false;
    if (!R.done) {
      if (R.rounds >= max_rounds) {
        fail_cleanup();
        CMIB_FAIL("the shoot did not finish in %llu rounds: packets are still queued (a re-emission probability of 1 "
                  "in a box that cannot be left?)", (unsigned long long)max_rounds);
      }
      /* without re-emission the rounds still needed are known (a shoot of more primaries than one group holds) */
      const int left = rounds_for_primaries(l) - (int)std::min<uint64_t>(R.rounds, 1u << 20);
      const int next = (!can_reemit && left > 0) ? std::min(left + 1, (int)CTL_STATUS_SLOTS - 1) : group;
      if (enqueue_group(l, next)) return 1;
      inflight.push_back(l);
    }
  }
  /* every group of every lane has been read back: nothing of this shoot is in flight */
  ctx->shoot_rounds = run[0].rounds + (nlanes == 2 ? run[1].rounds : 0);
  ctx->shoot_lanes = nlanes;
  {
    /* mean walk length of this shoot, for the key layout of the next one: the counters of the accumulator buffer
     * as the lane that finished last saw them */
    double c0, c1 = 0., e0, e1 = 0.;
    memcpy(&c0, &ctx->lane[0].h_ctl[((run[0].fin - 1) & 1) * CTL_WORDS + CTL_CROSSINGS0], 8);
    memcpy(&e0, &ctx->lane[0].h_ctl[((run[0].fin - 1) & 1) * CTL_WORDS + CTL_EMISSIONS0], 8);
    for (int l = 0; l < nlanes; ++l) {
      const unsigned long long *h = ctx->lane[l].h_ctl + ((run[l].fin - 1) & 1) * CTL_WORDS;
      double c, e;
      memcpy(&c, &h[CTL_CROSSINGS], 8); memcpy(&e, &h[CTL_EMISSIONS], 8);
      if (e > e1) { e1 = e; c1 = c; }
    }
    if (P.n_packets >= (1ull << 20) && e1 > e0) ctx->walk_cells = (c1 - c0) / (e1 - e0) / 1.5;
  }
  if (P.hot_replicas > 0) {
    const int n = P.src.n_sources * HOT_CELLS * HOT_STRIDE;
    if (mode == ACC_HONLY) fold_hot_cells_kernel<ACC_HONLY><<<blocks_for(n, 128), 128, 0, s0>>>(P);
    else fold_hot_cells_kernel<ACC_FULL><<<blocks_for(n, 128), 128, 0, s0>>>(P);
    ++g_launches;
    CUDA_OK(cudaGetLastError());
    CUDA_OK(cudaStreamSynchronize(s0));
  }
  ctx->prepare_ms = ctx->march_ms = ctx->overlap_ms = 0.;
  if (ctx->timing) {
    /* time with an emission kernel (decide, prepare) running, with a march kernel running, and with both
     * (two lanes): measures of unions of the kernels' intervals on a common clock (the origin event) */
    std::vector<std::pair<double, double>> prep, march;
    for (int l = 0; l < nlanes; ++l) {
      cmib_context::Lane &L = ctx->lane[l];
      for (size_t k = 0; k + 3 < L.ev_used; k += 4) {
        float t[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < 4; ++j) cudaEventElapsedTime(&t[j], ctx->origin_ev, L.ev_pool[k + j]);
        prep.emplace_back(t[0], t[1]);
        march.emplace_back(t[2], t[3]);
      }
    }
    merge_intervals(prep);
    merge_intervals(march);
    ctx->prepare_ms = interval_measure(prep);
    ctx->march_ms = interval_measure(march);
    ctx->overlap_ms = interval_overlap(prep, march);
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
  if ((double)grid->ncell[0] * grid->ncell[1] * grid->ncell[2] >= 4294967296.)
    CMIB_FAIL("grids of 2^32 cells or more are not supported (the walk carries 32-bit cell indices)");
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
  ctx->apply_conventions();
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
  for (cmib_context::Lane &L : ctx->lane) {
    if (L.h_ctl) cudaFreeHost(L.h_ctl);
    for (cudaEvent_t e : L.ev_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : L.group_done)
      if (e) cudaEventDestroy(e);
    if (L.stream && L.stream != ctx->stream) cudaStreamDestroy(L.stream);
  }
  for (cudaEvent_t e : {ctx->fork_ev, ctx->origin_ev})
    if (e) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->xev)
    if (e) cudaEventDestroy(e);
  if (ctx->comm) NcclApi::get().CommDestroy(ctx->comm);
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
  ctx->heat_possibly_written = false;
  return 0;
}

/* ---- plugins --------------------------------------------------------------- */
int cmib_set_abundances(cmib_context *ctx, const double *abundances) {
  CHECK_CTX(ctx);
  if (!abundances) CMIB_FAIL("null abundances");
  for (int k = 0; k < NUM_ELEMENTS; ++k) ctx->abund[k] = abundances[k];
  ctx->apply_conventions();
  return 0;
}

int cmib_set_packet_conventions(cmib_context *ctx, int conventions) {
  CHECK_CTX(ctx);
  if (conventions == CMIB_CONVENTIONS_IONIZATION_SIMULATION) {
    /* DensityGrid.hpp:219-222 */
    ctx->nu_H = ev_to_hz(13.6);
    ctx->nu_He = ev_to_hz(24.6);
  } else if (conventions == CMIB_CONVENTIONS_TASK_BASED) {
    /* DensitySubGrid.hpp:608-612 */
    ctx->nu_H = 3.288e15;
    ctx->nu_He = 5.948e15;
  } else {
    CMIB_FAIL("unknown packet conventions %d", conventions);
  }
  ctx->conventions = conventions;
  ctx->apply_conventions();
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
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
  }
  /* the replicas live at the end of the accumulator buffer: (re)size it for the new source list */
  if (ctx->acc.p && ensure_acc(ctx)) return 1;
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
    if (!(ctx->src.spectrum.kind == SPECTRUM_MONOCHROMATIC && ctx->src.spectrum.mono_frequency == ctx->nu_H &&
          ctx->src.reemission_kind == REEMISSION_NONE && ctx->src.continuous_kind == CONTINUOUS_NONE))
      ctx->heat_possibly_written = true;
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
    P.hot_acc = ctx->hot_acc();
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
  return cmib_update_state_block(ctx, loop, totweight, 0, (uint64_t)ctx->geom.ncells);
}

} /* extern "C" */

namespace {
/* number of work items of rank `rank` of `size` (kernels.cuh owned_cell): its chunks, the last one of the grid
 * counted in full; and the cells among them that exist */
int64_t owned_work_items(int64_t ncells, int32_t size, int32_t rank) {
  const int64_t nchunks = (ncells + OWN_CHUNK - 1) / OWN_CHUNK;
  return (nchunks > rank ? (nchunks - rank + size - 1) / size : 0) * OWN_CHUNK;
}
int64_t owned_cell_count(int64_t ncells, int32_t size, int32_t rank) {
  const int64_t nw = owned_work_items(ncells, size, rank);
  if (nw == 0) return 0;
  const int64_t last = owned_cell(nw - 1, rank, size); /* last cell of the rank's last chunk */
  return last < ncells ? nw : nw - (last + 1 - ncells);
}

/* the state update of the cells [cell_begin, cell_end) (own_size <= 1) or of the chunks rank own_rank of own_size owns */
int update_state_cells(cmib_context *ctx, uint32_t loop, double totweight, uint64_t cell_begin, uint64_t cell_end,
                       int32_t own_rank, int32_t own_size) {
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
  P.cell_begin = (int64_t)cell_begin;
  P.cell_end = (int64_t)cell_end;
  P.own_rank = own_rank;
  P.own_size = own_size;
  P.fold_abundances = (ctx->conventions == 1) ? 1 : 0;
  P.lc_wide_max_pairs = LC_WIDE_MAX_PAIRS;
  if (const char *e = getenv("CMIB_LC_WIDE")) P.lc_wide_max_pairs = atoi(e);
  P.n_work = (own_size > 1) ? owned_work_items(ctx->geom.ncells, own_size, own_rank) : (int64_t)(cell_end - cell_begin);
  const int64_t nc = P.n_work;
  if (nc == 0) return 0;
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
    CUDA_OK(cudaMemsetAsync(ctx->upd_counter.p, 0, sizeof(unsigned long long), ctx->stream)); /* next work item */
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
} // namespace

extern "C" {

int cmib_update_state_block(cmib_context *ctx, uint32_t loop, double totweight, uint64_t cell_begin, uint64_t cell_end) {
  CHECK_CTX(ctx);
  if (ensure_acc(ctx)) return 1;
  if (cell_end > (uint64_t)ctx->geom.ncells || cell_begin > cell_end) CMIB_FAIL("cell block outside the grid");
  if (cell_begin == cell_end) return 0;
  return update_state_cells(ctx, loop, totweight, cell_begin, cell_end, 0, 1);
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

int cmib_shoot_overlap(cmib_context *ctx, int32_t *lanes, double *overlap_ms) {
  CHECK_CTX(ctx);
  if (lanes) *lanes = ctx->shoot_lanes;
  if (overlap_ms) *overlap_ms = ctx->overlap_ms;
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
  if (n_doubles) *n_doubles = ctx->acc_main_doubles(ctx->acc_mode);
  return 0;
}

/* ---- multi-GPU: MPICommunicator of the reference for this path ----------------------------- */

uint64_t cmib_distribute(uint64_t number, int32_t size, int32_t rank) {
  /* MPICommunicator::distribute (MPICommunicator.hpp:207-222) */
  if (size <= 1) return number;
  const uint64_t quotient = number / (uint64_t)size, remainder = number % (uint64_t)size;
  return quotient + (((uint64_t)rank < remainder) ? 1u : 0u);
}

void cmib_distribute_block(int32_t rank, int32_t size, uint64_t begin, uint64_t end, uint64_t *block_begin,
                           uint64_t *block_end) {
  /* MPICommunicator::distribute_block (MPICommunicator.hpp:237-255) */
  const uint64_t block_size = end - begin;
  const uint64_t quotient = block_size / (uint64_t)size, remainder = block_size % (uint64_t)size;
  const uint64_t r = (uint64_t)rank;
  if (block_begin) *block_begin = begin + r * quotient + (r < remainder ? r : remainder);
  if (block_end) *block_end = begin + (r + 1) * quotient + (r + 1 < remainder ? r + 1 : remainder);
}

int cmib_comm_unique_id(void *id128) {
  if (!id128) CMIB_FAIL("null argument");
  NcclApi &nccl = NcclApi::get();
  if (!nccl.error.empty()) CMIB_FAIL("%s", nccl.error.c_str());
  static_assert(sizeof(ncclUniqueId) == 128, "the C ABI hands the NCCL unique id around as 128 bytes");
  ncclUniqueId id;
  NCCL_OK(nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int cmib_comm_init_rank(cmib_context *ctx, int32_t size, int32_t rank, const void *id128) {
  CHECK_CTX(ctx);
  if (!id128 || size < 1 || rank < 0 || rank >= size) CMIB_FAIL("bad communicator arguments (rank %d of %d)", rank, size);
  if (ctx->comm) CMIB_FAIL("the context already has a communicator");
  NcclApi &nccl = NcclApi::get();
  if (!nccl.error.empty()) CMIB_FAIL("%s", nccl.error.c_str());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  NCCL_OK(nccl.CommInitRank(&ctx->comm, size, id, rank));
  ctx->comm_rank = rank;
  ctx->comm_size = size;
  return 0;
}

int cmib_comm_init_all(cmib_context **ctxs, int32_t size) {
  if (!ctxs || size < 1) CMIB_FAIL("bad communicator arguments");
  NcclApi &nccl = NcclApi::get();
  if (!nccl.error.empty()) CMIB_FAIL("%s", nccl.error.c_str());
  std::vector<int> devices(size);
  for (int r = 0; r < size; ++r) {
    if (!ctxs[r]) CMIB_FAIL("null context");
    if (ctxs[r]->comm) CMIB_FAIL("context %d already has a communicator", r);
    devices[r] = ctxs[r]->device;
  }
  std::vector<ncclComm_t> comms(size);
  NCCL_OK(nccl.CommInitAll(comms.data(), size, devices.data()));
  for (int r = 0; r < size; ++r) {
    ctxs[r]->comm = comms[r];
    ctxs[r]->comm_rank = r;
    ctxs[r]->comm_size = size;
  }
  return 0;
}

int cmib_comm_finalize(cmib_context *ctx) {
  CHECK_CTX(ctx);
  if (ctx->comm) {
    CUDA_OK(cudaStreamSynchronize(ctx->stream));
    NcclApi::get().CommDestroy(ctx->comm);
    ctx->comm = nullptr;
  }
  ctx->comm_rank = 0;
  ctx->comm_size = 1;
  return 0;
}

int cmib_comm_info(cmib_context *ctx, int32_t *rank, int32_t *size, uint64_t *cell_begin, uint64_t *cell_end) {
  CHECK_CTX(ctx);
  if (rank) *rank = ctx->comm_rank;
  if (size) *size = ctx->comm_size;
  cmib_distribute_block(ctx->comm_rank, ctx->comm_size, 0, (uint64_t)ctx->geom.ncells, cell_begin, cell_end);
  return 0;
}

} /* extern "C" */

namespace {

/* all-gather of a per-cell array with `per_cell` elements of T per cell, blocks of distribute_block: one in-place
 * broadcast per rank inside a group (the blocks differ by at most one cell, so ncclAllGather's equal counts do not fit) */
template <typename T>
int gather_blocks(cmib_context *ctx, T *array, size_t per_cell) {
  NcclApi &nccl = NcclApi::get();
  NCCL_OK(nccl.GroupStart());
  for (int r = 0; r < ctx->comm_size; ++r) {
    uint64_t lo, hi;
    cmib_distribute_block(r, ctx->comm_size, 0, (uint64_t)ctx->geom.ncells, &lo, &hi);
    if (hi == lo) continue;
    T *blk = array + lo * per_cell;
    NCCL_OK(nccl.Broadcast(blk, blk, (hi - lo) * per_cell * sizeof(T), ncclChar, r, ctx->comm, ctx->stream));
  }
  NCCL_OK(nccl.GroupEnd());
  return 0;
}

/* all-gather of a per-cell array with `per_cell` doubles per cell whose chunks were updated by their owners
 * (kernels.cuh owned_cell): pack the own chunks, ncclAllGather of the equal-sized packs, scatter the others' */
int gather_owned(cmib_context *ctx, double *array, int per_cell, bool opacity_records) {
  NcclApi &nccl = NcclApi::get();
  const int64_t nc = ctx->geom.ncells;
  const int32_t size = ctx->comm_size, rank = ctx->comm_rank;
  const int64_t nw = owned_work_items(nc, size, 0); /* rank 0 owns the most chunks: the common pack size */
  const size_t pack = (size_t)nw * per_cell;
  if (ctx->xchg_pack.n < pack) CUDA_OK(ctx->xchg_pack.resize(pack));
  if (ctx->xchg_all.n < pack * size) CUDA_OK(ctx->xchg_all.resize(pack * size));
  cudaStream_t s = ctx->stream;
  if (opacity_records)
    pack_owned_records_kernel<<<blocks_for(nw, 256), 256, 0, s>>>(nw, nc, rank, size, reinterpret_cast<const CellOpacity *>(array),
                                                                reinterpret_cast<CellOpacity *>(ctx->xchg_pack.p));
  else
    pack_owned_doubles_kernel<<<blocks_for(nw * per_cell, 256), 256, 0, s>>>(nw, nc, rank, size, per_cell, array, ctx->xchg_pack.p);
  NCCL_OK(nccl.AllGather(ctx->xchg_pack.p, ctx->xchg_all.p, pack, ncclDouble, ctx->comm, s));
  if (opacity_records)
    unpack_owned_records_kernel<<<blocks_for(nw * size, 256), 256, 0, s>>>(nw, nc, rank, size, reinterpret_cast<const CellOpacity *>(ctx->xchg_all.p),
                                                                         ctx->cells.p, ctx->cells_h.p);
  else
    unpack_owned_doubles_kernel<<<blocks_for(nw * per_cell * size, 256), 256, 0, s>>>(nw, nc, rank, size, per_cell, ctx->xchg_all.p, array);
  g_launches += 2;
  CUDA_OK(cudaGetLastError());
  return 0;
}

} // namespace

extern "C" {

int cmib_comm_exchange_and_update(cmib_context *ctx, uint32_t loop, int allreduce) {
  CHECK_CTX(ctx);
  (void)allreduce; /* every rank receives all sums (see include/cmib.h) */
  if (ensure_acc(ctx)) return 1;
  if (!ctx->comm || ctx->comm_size == 1) return cmib_update_state(ctx, loop, 0.);
  cudaStream_t s = ctx->stream;
  if (!ctx->xev[0])
    for (int k = 0; k < 4; ++k) CUDA_OK(cudaEventCreate(&ctx->xev[k]));
  CUDA_OK(cudaEventRecord(ctx->xev[0], s));
  size_t count = ctx->acc_main_doubles(ctx->acc_mode);
  if (ctx->acc_mode == ACC_HONLY && ctx->honly_planar() && !ctx->heat_possibly_written && !getenv("CMIB_REDUCE_ALL"))
    count = ACC_COUNTERS + (size_t)ctx->honly_offset() + (size_t)ctx->geom.ncells; /* counters + the J_H plane; the heat plane is zero */
  NCCL_OK(NcclApi::get().AllReduce(ctx->acc.p, ctx->acc.p, count, ncclDouble, ncclSum, ctx->comm, s));
  CUDA_OK(cudaEventRecord(ctx->xev[1], s));
  /* totweight: the reduced counter on the device */
  if (update_state_cells(ctx, loop, 0., 0, (uint64_t)ctx->geom.ncells, ctx->comm_rank, ctx->comm_size)) return 1;
  CUDA_OK(cudaEventRecord(ctx->xev[2], s));
  static_assert(sizeof(CellOpacity) == 4 * sizeof(double), "opacity records travel as 4 doubles");
  if (gather_owned(ctx, reinterpret_cast<double *>(ctx->cells.p), 4, true)) return 1;
  CUDA_OK(cudaEventRecord(ctx->xev[3], s));
  ctx->reemit_prob_valid = false;
  return 0;
}

int cmib_comm_exchange_timing(cmib_context *ctx, double ms[3]) {
  CHECK_CTX(ctx);
  if (!ms) CMIB_FAIL("null argument");
  ms[0] = ms[1] = ms[2] = 0.;
  if (!ctx->xev[0]) return 0;
  CUDA_OK(cudaEventSynchronize(ctx->xev[3]));
  for (int k = 0; k < 3; ++k) {
    float f = 0.f;
    CUDA_OK(cudaEventElapsedTime(&f, ctx->xev[k], ctx->xev[k + 1]));
    ms[k] = f;
  }
  return 0;
}

int cmib_comm_gather_state(cmib_context *ctx) {
  CHECK_CTX(ctx);
  if (!ctx->comm || ctx->comm_size == 1) return 0;
  if (gather_owned(ctx, ctx->xmetal.p, 12, false)) return 1;
  if (gather_owned(ctx, ctx->heat_norm.p, 2, false)) return 1;
  return 0;
}

int cmib_comm_owned_cells(cmib_context *ctx, uint64_t *n_owned) {
  CHECK_CTX(ctx);
  if (!n_owned) CMIB_FAIL("null argument");
  *n_owned = (ctx->comm_size > 1) ? (uint64_t)owned_cell_count(ctx->geom.ncells, ctx->comm_size, ctx->comm_rank) : (uint64_t)ctx->geom.ncells;
  return 0;
}

uint64_t cmib_owned_cell(uint64_t j, int32_t size, int32_t rank) {
  return size > 1 ? (uint64_t)owned_cell((int64_t)j, rank, size) : j;
}

uint64_t cmib_owned_cell_count(uint64_t ncells, int32_t size, int32_t rank) {
  return size > 1 ? (uint64_t)owned_cell_count((int64_t)ncells, size, rank) : ncells;
}

int cmib_upload_cells_owned(cmib_context *ctx, const double *n, const double *T, const double *x) {
  CHECK_CTX(ctx);
  if (ctx->comm_size <= 1) return cmib_upload_cells(ctx, n, T, x, nullptr);
  if (!n || !T || !x) CMIB_FAIL("null cell array");
  const size_t no = (size_t)owned_cell_count(ctx->geom.ncells, ctx->comm_size, ctx->comm_rank);
  if (no == 0) return 0;
  if (ctx->stage.n < no * 18) CUDA_OK(ctx->stage.resize(no * 18));
  double *s = ctx->stage.p;
  CUDA_OK(cudaMemcpyAsync(s, n, no * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(s + no, T, no * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(s + 2 * no, x, no * 14 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  pack_cells_owned_kernel<<<blocks_for(no, 256), 256, 0, ctx->stream>>>((int64_t)no, ctx->comm_rank, ctx->comm_size, s, s + no, s + 2 * no,
                                                                        ctx->cells.p, ctx->cells_h.p, ctx->xmetal.p);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  ctx->reemit_prob_valid = false;
  CUDA_OK(cudaStreamSynchronize(ctx->stream)); /* host arrays may be reused on return */
  return 0;
}

int cmib_comm_gather_owned_cells(cmib_context *ctx) {
  CHECK_CTX(ctx);
  if (!ctx->comm || ctx->comm_size == 1) return 0;
  if (gather_owned(ctx, reinterpret_cast<double *>(ctx->cells.p), 4, true)) return 1;
  ctx->reemit_prob_valid = false;
  return 0;
}

int cmib_download_cells_owned(cmib_context *ctx, double *n, double *T, double *x, double *heat) {
  CHECK_CTX(ctx);
  if (ctx->comm_size <= 1) return cmib_download_cells(ctx, n, T, x, heat);
  const size_t no = (size_t)owned_cell_count(ctx->geom.ncells, ctx->comm_size, ctx->comm_rank);
  if (no == 0) return 0;
  if (ctx->stage.n < no * 18) CUDA_OK(ctx->stage.resize(no * 18));
  double *s = ctx->stage.p;
  unpack_cells_owned_kernel<<<blocks_for(no, 256), 256, 0, ctx->stream>>>((int64_t)no, ctx->comm_rank, ctx->comm_size, ctx->cells.p,
                                                                          ctx->xmetal.p, ctx->heat_norm.p, s, s + no, s + 2 * no, s + 16 * no);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  if (n) CUDA_OK(cudaMemcpyAsync(n, s, no * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (T) CUDA_OK(cudaMemcpyAsync(T, s + no, no * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (x) CUDA_OK(cudaMemcpyAsync(x, s + 2 * no, no * 14 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (heat) CUDA_OK(cudaMemcpyAsync(heat, s + 16 * no, no * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

int cmib_comm_gather_cells_all(cmib_context *ctx) {
  CHECK_CTX(ctx);
  if (!ctx->comm || ctx->comm_size == 1) return 0;
  if (gather_blocks(ctx, ctx->cells.p, 1)) return 1;
  if (gather_blocks(ctx, ctx->cells_h.p, 1)) return 1;
  if (gather_blocks(ctx, ctx->xmetal.p, 12)) return 1;
  ctx->reemit_prob_valid = false;
  return 0;
}

int cmib_upload_cells_block(cmib_context *ctx, uint64_t cell_begin, uint64_t cell_end, const double *n, const double *T,
                            const double *x) {
  CHECK_CTX(ctx);
  if (!n || !T || !x) CMIB_FAIL("null cell array");
  if (cell_end > (uint64_t)ctx->geom.ncells || cell_begin > cell_end) CMIB_FAIL("cell block outside the grid");
  const size_t nb = (size_t)(cell_end - cell_begin);
  if (nb == 0) return 0;
  if (ctx->stage.n < nb * 18) CUDA_OK(ctx->stage.resize(nb * 18));
  double *s = ctx->stage.p;
  CUDA_OK(cudaMemcpyAsync(s, n, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(s + nb, T, nb * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  CUDA_OK(cudaMemcpyAsync(s + 2 * nb, x, nb * 14 * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
  pack_cells_kernel<<<blocks_for(nb, 256), 256, 0, ctx->stream>>>((int64_t)nb, s, s + nb, s + 2 * nb, ctx->cells.p + cell_begin,
                                                                  ctx->cells_h.p + cell_begin, ctx->xmetal.p + cell_begin * 12);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  ctx->reemit_prob_valid = false;
  CUDA_OK(cudaStreamSynchronize(ctx->stream)); /* host arrays may be reused on return */
  return 0;
}

int cmib_download_cells_block(cmib_context *ctx, uint64_t cell_begin, uint64_t cell_end, double *n, double *T, double *x,
                              double *heat) {
  CHECK_CTX(ctx);
  if (cell_end > (uint64_t)ctx->geom.ncells || cell_begin > cell_end) CMIB_FAIL("cell block outside the grid");
  const size_t nb = (size_t)(cell_end - cell_begin);
  if (nb == 0) return 0;
  if (ctx->stage.n < nb * 18) CUDA_OK(ctx->stage.resize(nb * 18));
  double *s = ctx->stage.p;
  unpack_cells_kernel<<<blocks_for(nb, 256), 256, 0, ctx->stream>>>((int64_t)nb, ctx->cells.p + cell_begin,
                                                                    ctx->xmetal.p + cell_begin * 12,
                                                                    ctx->heat_norm.p + cell_begin * 2, s, s + nb, s + 2 * nb,
                                                                    s + 16 * nb);
  ++g_launches;
  CUDA_OK(cudaGetLastError());
  if (n) CUDA_OK(cudaMemcpyAsync(n, s, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (T) CUDA_OK(cudaMemcpyAsync(T, s + nb, nb * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (x) CUDA_OK(cudaMemcpyAsync(x, s + 2 * nb, nb * 14 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  if (heat) CUDA_OK(cudaMemcpyAsync(heat, s + 16 * nb, nb * 2 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CUDA_OK(cudaStreamSynchronize(ctx->stream));
  return 0;
}

/* ---- measured ceilings of the part ------------------------------------------------------------ */
int cmib_measure_scatter_rates(cmib_context *ctx, uint64_t n_cells, double *red_per_s, double *gather_per_s) {
  CHECK_CTX(ctx);
  if (n_cells == 0) CMIB_FAIL("empty table");
  DevBuf<double> table;
  DevBuf<double> out;
  CUDA_OK(table.resize((size_t)n_cells * 2)); /* 16-byte records, as the H-only walk gathers them */
  CUDA_OK(out.resize(1));
  cudaStream_t s = ctx->stream;
  CUDA_OK(cudaMemsetAsync(table.p, 0, (size_t)n_cells * 16, s));
  const int bs = 256, grid = ctx->sm_count * 8, iters = 64;
  cudaEvent_t e0, e1;
  CUDA_OK(cudaEventCreate(&e0));
  CUDA_OK(cudaEventCreate(&e1));
  double best[2] = {1e30, 1e30};
  for (int rep = 0; rep < 4; ++rep) { /* the first repetition is the warm-up */
    for (int which = 0; which < 2; ++which) {
      CUDA_OK(cudaEventRecord(e0, s));
      if (which == 0) measure_scatter_red_kernel<<<grid, bs, 0, s>>>(table.p, n_cells, iters);
      else measure_scatter_gather_kernel<<<grid, bs, 0, s>>>(reinterpret_cast<const double2 *>(table.p), n_cells, iters, out.p);
      CUDA_OK(cudaEventRecord(e1, s));
      CUDA_OK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
      if (rep > 0 && ms < best[which]) best[which] = ms;
    }
  }
  g_launches += 8;
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  table.resize(0);
  out.resize(0);
  const double ops = (double)bs * grid * iters;
  if (red_per_s) *red_per_s = ops / (best[0] * 1e-3);
  if (gather_per_s) *gather_per_s = ops / (best[1] * 1e-3);
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
  ctx->heat_possibly_written = true; /* explicit packets carry any frequency */
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
