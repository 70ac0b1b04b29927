/*
 * wavefront.cuh — the production shoot path: photon packets flow through two
 * warp-convergent kernels connected by device-resident queues.
 *
 *   prepare_kernel   (re-)emission: decides the fate of absorbed packets
 *                    (PhotonSource::reemit, src/PhotonSource.cpp:272-308), draws fresh
 *                    primaries (PhotonSource::get_random_photon, :208-249), samples the
 *                    new direction / frequency / 14 cross sections / optical depth and
 *                    appends the packet to the march queue.  All lanes of a warp
 *                    execute the pow()-heavy code together.
 *   march_kernel     CartesianDensityGrid::interact (src/CartesianDensityGrid.cpp:375-452)
 *                    + DensityGrid::update_integrals (src/DensityGrid.hpp:150-197) as a
 *                    persistent warp state machine: every pass all live lanes take
 *                    exactly one cell crossing; lanes whose packet ended are refilled
 *                    from the queue in batches (warp-level compaction of live packets);
 *                    packets absorbed inside the box are appended to the re-emission
 *                    queue with one warp-aggregated atomic.
 *
 * Why: in a one-thread-per-packet kernel (shoot_kernel, kept as the A/B check)
 * ncu showed 3.9 active threads per warp instruction: lanes sit in different
 * phases (emission vs walk) and the 46 pow() of the Verner fits run one lane at a
 * time (profiles/r01_shoot_simple.md).  Here each phase is its own kernel and the
 * per-packet RNG stream (rng.cuh) is carried through the queues, so the packets,
 * their trajectories and therefore all sums are those of shoot_packet (shoot.cuh)
 * up to the order of the atomic adds.
 *
 * Queue entries are structure-of-arrays with the queue capacity as the stride.
 */
#pragma once
#include "cmib_common.cuh"
#include "march.cuh"
#include "rng.cuh"
#include "shoot.cuh"
#include "source.cuh"

namespace cmib {

/* control block (uint64 words) */
enum CtlWord : int {
  CTL_QCOUNT = 0,   /* entries in the march queue */
  CTL_RQCOUNT,      /* entries in the re-emission queue */
  CTL_HEAD,         /* next unclaimed march-queue entry */
  CTL_REMAINING,    /* primaries not yet emitted */
  CTL_NEXT_FRESH,   /* local index of the next primary */
  CTL_ROUND,
  CTL_STATUS = 8,   /* ring of CTL_STATUS_SLOTS words: march-queue size after each prepare */
  CTL_STATUS_SLOTS = 64,
  CTL_WORDS = CTL_STATUS + CTL_STATUS_SLOTS
};

/* march-queue fields (8-byte each) */
enum MarchField : int { MQ_PX = 0, MQ_PY, MQ_PZ, MQ_DX, MQ_DY, MQ_DZ, MQ_NU, MQ_TAU, MQ_ID, MQ_META, MQ_SIGMA };
template <int MODE> struct MarchQueueLayout {
  /* sigma[NSIG], then A_He*sigma_He in the full layout */
  static constexpr int NFIELDS = MQ_SIGMA + AccLayout<MODE>::NSIG + (MODE == ACC_FULL ? 1 : 0);
};
/* re-emission-queue fields */
enum ReemitField : int { RQ_PX = 0, RQ_PY, RQ_PZ, RQ_SIGH, RQ_SIGHE, RQ_CELL, RQ_ID, RQ_META, RQ_NFIELDS };

struct WavefrontParams {
  ShootParams sp;
  unsigned long long *ctl;
  double *mq;              /* march queue: [NFIELDS][capacity] */
  double *rq;              /* re-emission queue: [RQ_NFIELDS][capacity] */
  uint64_t capacity;
};

CMIB_D uint64_t pack_meta(uint32_t ndraw, int type) { return ((uint64_t)(uint32_t)type << 32) | ndraw; }

/* number of uniforms consumed so far */
CMIB_D uint32_t rng_save(const PacketRng &r) { return 2u * r.block - r.have; }

CMIB_D void rng_restore(PacketRng &r, uint64_t seed, uint32_t iteration, uint64_t packet_id, uint32_t ndraw) {
  rng_init(r, seed, iteration, packet_id);
  if (ndraw & 1u) {
    r.block = ndraw >> 1;
    (void)rng_uniform(r); /* regenerates the block; leaves its second half in `spare` */
  } else {
    r.block = ndraw >> 1;
  }
}

/* block-wide sum of per-thread counters into the 7 leading doubles of acc */
CMIB_D void reduce_counters(double *acc, const ShootCounters &cnt) {
  __shared__ double red[7][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[7] = {cnt.w_tot, cnt.w_type[0], cnt.w_type[1], cnt.w_type[2], cnt.w_type[3],
                 (double)cnt.n_steps, (double)cnt.n_emit};
#pragma unroll
  for (int k = 0; k < 7; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double sum = 0.;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) sum += red[threadIdx.x][w];
    if (sum != 0.) atomicAdd(acc + threadIdx.x, sum);
  }
}

/* ------------------------------------------------------------------------- */
/* prepare: re-emission decisions + fresh primaries -> march queue            */
/* ------------------------------------------------------------------------- */
template <int MODE>
__global__ void __launch_bounds__(256)
prepare_kernel(const __grid_constant__ WavefrontParams W) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  const ShootParams &P = W.sp;
  const SourceModel &m = P.src;
  const uint64_t cap = W.capacity;
  const uint64_t n_re = W.ctl[CTL_RQCOUNT];
  const uint64_t remaining = W.ctl[CTL_REMAINING];
  const uint64_t next_fresh = W.ctl[CTL_NEXT_FRESH];
  const uint64_t room = cap - n_re;
  const uint64_t n_fresh = remaining < room ? remaining : room;
  const uint64_t n_items = n_re + n_fresh;
  ShootCounters cnt;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  /* warp-uniform trip count so that the ballot below is executed by whole warps */
  const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
  for (uint64_t w0 = first; w0 < n_items; w0 += stride) {
    const uint64_t w = w0 + lane;
    bool emit = false;
    PacketRng rng;
    double px = 0., py = 0., pz = 0., nu = 0.;
    uint64_t id = 0;
    int type = PACKET_PRIMARY;
    if (w < n_re) {
      /* --- PhotonSource::reemit --- */
      px = W.rq[RQ_PX * cap + w];
      py = W.rq[RQ_PY * cap + w];
      pz = W.rq[RQ_PZ * cap + w];
      const double sigH = W.rq[RQ_SIGH * cap + w];
      const double sigHe = W.rq[RQ_SIGHE * cap + w];
      const int64_t cell = __double_as_longlong(W.rq[RQ_CELL * cap + w]);
      id = (uint64_t)__double_as_longlong(W.rq[RQ_ID * cap + w]);
      const uint64_t meta = (uint64_t)__double_as_longlong(W.rq[RQ_META * cap + w]);
      rng_restore(rng, P.seed, P.iteration, id, (uint32_t)meta);
      type = (int)(meta >> 32);
      if (m.reemission_kind == REEMISSION_PHYSICAL) {
        const CellOpacity c = load_cell(P.cells, cell);
        double p[NUM_REEMIT];
#pragma unroll
        for (int k = 0; k < NUM_REEMIT; ++k) p[k] = P.reemit_prob[cell * NUM_REEMIT + k];
        nu = physical_reemit(m, sigH, sigHe, c.xH, c.xHe, c.T, p, rng, type);
      } else { /* REEMISSION_FIXED (REEMISSION_NONE never queues) */
        const double u = rng_uniform(rng);
        if (u < m.fixed_reemission_probability) {
          type = PACKET_DIFFUSE_HI;
          nu = m.fixed_reemission_frequency;
        } else {
          type = PACKET_ABSORBED;
        }
      }
      if (nu == 0.) {
        /* absorbed for good: IonizationPhotonShootJob.hpp:143-144 */
        cnt.w_tot += m.discrete_weight;
#pragma unroll
        for (int t = 0; t < NUM_PACKET_TYPES; ++t) cnt.w_type[t] += (t == type) ? m.discrete_weight : 0.;
      } else {
        emit = true;
      }
    } else if (w < n_items) {
      /* --- PhotonSource::get_random_photon --- */
      id = P.packet_offset + next_fresh + (w - n_re);
      rng_init(rng, P.seed, P.iteration, id);
      double x = rng_uniform(rng);
      (void)x; /* discrete vs continuous draw: consumed as in the reference */
      x = rng_uniform(rng);
      int isrc = 0;
      while (isrc < m.n_sources - 1 && x > m.src_cum[isrc]) ++isrc;
      px = m.src_pos[3 * isrc];
      py = m.src_pos[3 * isrc + 1];
      pz = m.src_pos[3 * isrc + 2];
      emit = true;
    }
    double dx = 0., dy = 0., dz = 0., tau = 0., sigma_He_corr = 0.;
    double sigma[NSIG];
    if (emit) {
      ++cnt.n_emit;
      random_direction(rng, dx, dy, dz);
      if (w >= n_re) nu = (m.spectrum_kind == SPECTRUM_PLANCK) ? planck_frequency(m.planck, rng) : m.mono_frequency;
      packet_cross_sections<NSIG>(m, nu, sigma, sigma_He_corr);
      tau = -log(rng_uniform(rng));
    }
    /* warp-aggregated append */
    const unsigned ballot = __ballot_sync(0xffffffffu, emit);
    if (ballot) {
      unsigned long long base = 0;
      if (lane == (__ffs(ballot) - 1)) base = atomicAdd(&W.ctl[CTL_QCOUNT], (unsigned long long)__popc(ballot));
      base = __shfl_sync(0xffffffffu, base, __ffs(ballot) - 1);
      if (emit) {
        const uint64_t slot = base + __popc(ballot & ((1u << lane) - 1u));
        double *q = W.mq + slot;
        q[MQ_PX * cap] = px; q[MQ_PY * cap] = py; q[MQ_PZ * cap] = pz;
        q[MQ_DX * cap] = dx; q[MQ_DY * cap] = dy; q[MQ_DZ * cap] = dz;
        q[MQ_NU * cap] = nu; q[MQ_TAU * cap] = tau;
        q[MQ_ID * cap] = __longlong_as_double((long long)id);
        q[MQ_META * cap] = __longlong_as_double((long long)pack_meta(rng_save(rng), type));
#pragma unroll
        for (int k = 0; k < NSIG; ++k) q[(MQ_SIGMA + k) * cap] = sigma[k];
        if (MODE == ACC_FULL) q[(MQ_SIGMA + NSIG) * cap] = sigma_He_corr;
      }
    }
  }
  reduce_counters(P.acc, cnt);
}

/* bookkeeping after prepare: consume the primaries, empty the re-emission queue */
__global__ void advance_after_prepare_kernel(unsigned long long *ctl, uint64_t capacity) {
  const uint64_t n_re = ctl[CTL_RQCOUNT];
  const uint64_t remaining = ctl[CTL_REMAINING];
  const uint64_t room = capacity - n_re;
  const uint64_t n_fresh = remaining < room ? remaining : room;
  ctl[CTL_REMAINING] = remaining - n_fresh;
  ctl[CTL_NEXT_FRESH] += n_fresh;
  ctl[CTL_RQCOUNT] = 0;
  ctl[CTL_HEAD] = 0;
  const uint64_t round = ctl[CTL_ROUND];
  ctl[CTL_STATUS + (round % CTL_STATUS_SLOTS)] = ctl[CTL_QCOUNT];
  ctl[CTL_ROUND] = round + 1;
}

__global__ void advance_after_march_kernel(unsigned long long *ctl) { ctl[CTL_QCOUNT] = 0; }

/* ------------------------------------------------------------------------- */
/* march: persistent warp state machine over the march queue                  */
/* ------------------------------------------------------------------------- */
constexpr int MARCH_BLOCK = 256;
constexpr int MARCH_CHUNK = 128;     /* queue entries a warp claims with one atomic */
constexpr int MARCH_REFILL_MIN = 8;  /* idle lanes that trigger a refill while the queue has entries */

template <int MODE>
__global__ void __launch_bounds__(MARCH_BLOCK)
march_kernel(const __grid_constant__ WavefrontParams W) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  constexpr int NMETAL = (MODE == ACC_FULL) ? 12 : 0;
  __shared__ double s_sig[(NMETAL > 0 ? NMETAL : 1)][MARCH_BLOCK];
  const ShootParams &P = W.sp;
  const GridGeom &g = P.geom;
  const uint64_t cap = W.capacity;
  const uint64_t qcount = W.ctl[CTL_QCOUNT];
  const int lane = threadIdx.x & 31;
  const bool can_reemit = (P.src.reemission_kind != REEMISSION_NONE);
  const double weight = P.src.discrete_weight;
  ShootCounters cnt;

  MarchState s;
  double sigH = 0., sigHe = 0., sigHe_corr = 0., dnu_H = 0., dnu_He = 0.;
  uint64_t id = 0, meta = 0;
  uint32_t mask = 0;   /* metals (bits 2..13) with a non-zero cross section */
  bool has = false, live = false;
  s.last_cell = -1;
  /* warp-uniform cursor into the claimed chunk */
  uint64_t cur = 0, end = 0;
  bool exhausted = (qcount == 0);

  while (true) {
    /* ---- refill: hand queue entries to idle lanes ---- */
    const unsigned idle = __ballot_sync(0xffffffffu, !has);
    const int nidle = __popc(idle);
    if (nidle > 0 && !exhausted && (nidle >= MARCH_REFILL_MIN || idle == 0xffffffffu)) {
      if (cur == end) {
        unsigned long long b = 0;
        if (lane == 0) b = atomicAdd(&W.ctl[CTL_HEAD], (unsigned long long)MARCH_CHUNK);
        b = __shfl_sync(0xffffffffu, b, 0);
        if (b >= qcount) {
          exhausted = true;
        } else {
          cur = b;
          end = (b + MARCH_CHUNK < qcount) ? b + MARCH_CHUNK : qcount;
        }
      }
      if (!exhausted) {
        const int rank = __popc(idle & ((1u << lane) - 1u));
        const uint64_t avail = end - cur;
        if (!has && (uint64_t)rank < avail) {
          const double *q = W.mq + (cur + rank);
          s.px = q[MQ_PX * cap]; s.py = q[MQ_PY * cap]; s.pz = q[MQ_PZ * cap];
          s.dx = q[MQ_DX * cap]; s.dy = q[MQ_DY * cap]; s.dz = q[MQ_DZ * cap];
          const double nu = q[MQ_NU * cap];
          s.tau = q[MQ_TAU * cap];
          id = (uint64_t)__double_as_longlong(q[MQ_ID * cap]);
          meta = (uint64_t)__double_as_longlong(q[MQ_META * cap]);
          sigH = q[MQ_SIGMA * cap];
          mask = 0;
          if (MODE == ACC_FULL) {
            sigHe = q[(MQ_SIGMA + 1) * cap];
            sigHe_corr = q[(MQ_SIGMA + NSIG) * cap];
#pragma unroll
            for (int k = 0; k < NMETAL; ++k) {
              const double v = q[(MQ_SIGMA + 2 + k) * cap];
              s_sig[k][threadIdx.x] = v;
              mask |= (v != 0.) ? (1u << (2 + k)) : 0u;
            }
          }
          dnu_H = nu - P.nu_H;
          dnu_He = nu - P.nu_He;
          s.ix_ = 1. / s.dx;
          s.iy_ = 1. / s.dy;
          s.iz_ = 1. / s.dz;
          march_locate(g, s);
          has = true;
          live = march_inside(g, s) && s.tau > 0.;
        }
        cur += ((uint64_t)nidle < avail) ? (uint64_t)nidle : avail;
      }
    }
    if (__ballot_sync(0xffffffffu, has) == 0u) {
      if (exhausted) break;
      continue;
    }

    /* ---- one cell crossing for every live lane ---- */
    if (has && live) {
      const int64_t cell = long_index(g, s.ix, s.iy, s.iz);
      s.last_cell = cell;
      const CellOpacity c = load_cell(P.cells, cell);
      const double ds = march_step(g, s, c.n, c.xH, c.xHe, sigH, sigHe_corr);
      if (c.n > 0.) {
        /* update_integrals (DensityGrid.hpp:150-197); zero increments are skipped (exact) */
        const double dsw = ds * weight;
        double *a = P.acc + ACC_COUNTERS + cell * AccLayout<MODE>::NACC;
        const double dJH = dsw * sigH;
        if (dJH != 0.) {
          atomicAdd(a + ION_H_n, dJH);
          const double dh = dJH * dnu_H;
          if (dh != 0.) atomicAdd(a + (MODE == ACC_FULL ? NUM_IONS + HEAT_H : 1), dh);
        }
        if (MODE == ACC_FULL) {
          const double dJHe = dsw * sigHe;
          if (dJHe != 0.) {
            atomicAdd(a + ION_He_n, dJHe);
            const double dh = dJHe * dnu_He;
            if (dh != 0.) atomicAdd(a + NUM_IONS + HEAT_He, dh);
          }
          uint32_t mm = mask;
          while (mm) {
            const int k = __ffs(mm) - 1;
            mm &= mm - 1u;
            const double dJ = dsw * s_sig[k - 2][threadIdx.x];
            if (dJ != 0.) atomicAdd(a + k, dJ);
          }
        }
      }
      ++cnt.n_steps;
      live = march_inside(g, s) && s.tau > 0.;
    }

    /* ---- packets that ended: escaped (left the box) or absorbed (tau used up inside) ---- */
    const bool fin = has && !live;
    /* march_inside is idempotent once the periodic wrap has been applied */
    const bool inside = fin && march_inside(g, s);
    const bool absorbed = fin && inside;
    if (fin && !(absorbed && can_reemit)) {
      int type = (int)(meta >> 32);
      if (absorbed) type = PACKET_ABSORBED; /* PhotonSource::reemit without a handler (:304-306) */
      cnt.w_tot += weight;
#pragma unroll
      for (int t = 0; t < NUM_PACKET_TYPES; ++t) cnt.w_type[t] += (t == type) ? weight : 0.;
    }
    if (can_reemit) {
      const unsigned ab = __ballot_sync(0xffffffffu, absorbed);
      if (ab) {
        unsigned long long base = 0;
        const int leader = __ffs(ab) - 1;
        if (lane == leader) base = atomicAdd(&W.ctl[CTL_RQCOUNT], (unsigned long long)__popc(ab));
        base = __shfl_sync(0xffffffffu, base, leader);
        if (absorbed) {
          double *q = W.rq + (base + __popc(ab & ((1u << lane) - 1u)));
          q[RQ_PX * cap] = s.px; q[RQ_PY * cap] = s.py; q[RQ_PZ * cap] = s.pz;
          q[RQ_SIGH * cap] = sigH;
          q[RQ_SIGHE * cap] = sigHe;
          q[RQ_CELL * cap] = __longlong_as_double((long long)s.last_cell);
          q[RQ_ID * cap] = __longlong_as_double((long long)id);
          q[RQ_META * cap] = __longlong_as_double((long long)meta);
        }
      }
    }
    if (fin) has = false;
  }
  reduce_counters(P.acc, cnt);
}

} // namespace cmib
