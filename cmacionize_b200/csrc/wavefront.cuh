/*
 * wavefront.cuh — the production shoot path: photon packets flow through three
 * warp-convergent kernels connected by device-resident queues.
 *
 *   reemit_decide_kernel  fate of the packets absorbed in the last round (PhotonSource::reemit,
 *                    src/PhotonSource.cpp:272-308 up to the new frequency); survivors are
 *                    compacted into the emission queue.
 *   prepare_kernel   emission-queue entries + fresh primaries (PhotonSource::get_random_photon,
 *                    :208-249): new direction, frequency, 14 cross sections, optical depth;
 *                    item w -> march-queue slot w.  All lanes of a warp execute the
 *                    exp/log-heavy code together.
 *   march_kernel     CartesianDensityGrid::interact (src/CartesianDensityGrid.cpp:375-452)
 *                    + DensityGrid::update_integrals (src/DensityGrid.hpp:150-197) as a
 *                    persistent warp state machine: every pass all live lanes take
 *                    exactly one cell crossing; lanes whose packet ended are finished and
 *                    refilled from the queue in batches (warp-level compaction of live
 *                    packets); packets absorbed inside the box are appended to the
 *                    re-emission queue with one warp-aggregated atomic.
 *   fold_hot_cells_kernel  adds the replicated accumulators of the cells around the sources.
 *   sort_*_kernel    counting sort of the march queue for the coherent march (grids that do not fit in L2).
 *   tail_kernel      the last generations of re-emitted packets in one launch (packet_walks of shoot.cuh).
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
 * Host side: shoot_wavefront() in cmib_api.cu (rounds of decide -> prepare -> march).
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
  CTL_ERROR,        /* set by a kernel that hit its safety valve */
  CTL_EQCOUNT,      /* entries in the emission queue (re-emissions that survived the decision) */
  CTL_STATUS = 8,   /* ring of CTL_STATUS_SLOTS words: march-queue size after each prepare */
  CTL_STATUS_SLOTS = 64,
  /* walk statistics of this shoot (bit patterns of doubles): the counters of the accumulator buffer when the shoot
   * began and after the last march, so that the host learns the mean walk length with the read-back it does anyway */
  CTL_CROSSINGS0 = CTL_STATUS + CTL_STATUS_SLOTS, CTL_EMISSIONS0, CTL_CROSSINGS, CTL_EMISSIONS,
  CTL_TAIL_BLOCKS,  /* CTAs of tail_kernel that have finished (returns to 0 by itself) */
  CTL_WORDS
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
  double *eq;              /* emission queue: [EQ_NFIELDS][capacity] */
  uint64_t capacity;       /* stride of the queue arrays */
  uint64_t fill;           /* entries one round may hold (<= capacity; set per round by the host) */
  /* coherent march (sort == 2): the march queue is read in the order of a key (source | direction of a primary,
   * start position of a re-emitted packet); counting sort: prepare_kernel takes a ticket in its key's bin,
   * sort_scan_* turn the bin counts into offsets, sort_scatter_kernel writes the order */
  int sort;                /* 0: march reads the queue in emission order, 2: in key order */
  uint32_t *key;           /* [capacity] key of each queue entry */
  uint32_t *rank;          /* [capacity] ticket of each queue entry inside its bin */
  uint32_t *order;         /* [capacity] queue entries in key order */
  uint32_t *hist;          /* [nbins] entries per bin (zero between rounds) */
  uint32_t *offs;          /* [nbins] first position of each bin in the order */
  uint32_t *block_sums;    /* [nbins / SORT_SCAN_TILE] */
  uint32_t nbins;          /* 2^(fine_key_bits + 1), a multiple of SORT_SCAN_TILE */
  int fine_dir_bits;       /* direction bits of a primary's key (even, <= 22) */
  int fine_key_bits;       /* source + direction + optical-depth bits; bit fine_key_bits flags a re-emitted packet */
  int tau_bits;            /* low bits of a key: bin of the sampled optical depth (uniform in exp(-tau)) */
  uint32_t chunk_stride;   /* chunk c of the ordered queue is claimed as (c * chunk_stride) % nchunks */
  int agg;                 /* 1: march_kernel<MODE, true> (in-warp sums), 0: the plain kernel on the ordered queue */
  uint32_t lean_n16[3];    /* march_lean_kernel: 16 * ncell per axis */
  int32_t lean_k[2];       /* ... and the cell-index strides ncy * ncz, ncz */
  double uni_sigH, uni_w;  /* march_lean_kernel<HEAT = false>: sigma_H and weight of every packet of the shoot */
  uint32_t hot_index0;     /* first hot-cell replica record, in doubles from acc_j (the replicas follow the accumulators
                            * in the same allocation) */
  double *acc_j;           /* H-only layout: accumulator J_H of cell 0 (sp.acc + ACC_COUNTERS + sp.honly_offset) */
};

/*
 * Keys of the coherent march (sort == 2).  Packets that start at the same source and leave in nearly the
 * same direction walk through the same cells: queue entries are ordered by key so that the packets in flight
 * at any time — the warps claim consecutive chunks of the ordered queue — cover a narrow cone of the grid that
 * stays in L2, and the 8 lanes of a refill group cross the same cells in lock-step.  Keys are
 * fine_key_bits + 1 bits wide (<= 23: the bin tables stay L2 resident): primaries  0 | source | direction on a
 * 2048 x 2048 octahedral map in Morton order (truncated to fine_dir_bits); re-emitted packets  1 | 1024^3
 * Morton bin of the start position (truncated).  With 1.6e7 packets of one source in the queue a bin holds ~4
 * packets that leave within ~0.002 rad of each other.  Ordering changes which packets run together, not what
 * any packet does: results are identical up to the order of the atomic adds.
 */
CMIB_D uint32_t spread2(uint32_t x) { /* 16 bits -> even bit positions */
  x &= 0xffffu;
  x = (x | (x << 8)) & 0x00ff00ffu;
  x = (x | (x << 4)) & 0x0f0f0f0fu;
  x = (x | (x << 2)) & 0x33333333u;
  x = (x | (x << 1)) & 0x55555555u;
  return x;
}
CMIB_D uint32_t spread3(uint32_t x) { /* 10 bits -> every third bit position */
  x &= 0x3ffu;
  x = (x | (x << 16)) & 0x030000ffu;
  x = (x | (x << 8)) & 0x0300f00fu;
  x = (x | (x << 4)) & 0x030c30c3u;
  x = (x | (x << 2)) & 0x09249249u;
  return x;
}
CMIB_D uint32_t fine_direction_key(double dx, double dy, double dz) { /* 22 bits */
  const double ax = fabs(dx), ay = fabs(dy), az = fabs(dz);
  const double inv = 1. / fmax(ax + ay + az, 1e-300);
  double u = dx * inv, v = dy * inv;
  if (dz < 0.) {
    const double uu = (1. - fabs(v)) * (u >= 0. ? 1. : -1.);
    const double vv = (1. - fabs(u)) * (v >= 0. ? 1. : -1.);
    u = uu; v = vv;
  }
  const int iu = min(2047, max(0, (int)((u * 0.5 + 0.5) * 2048.)));
  const int iv = min(2047, max(0, (int)((v * 0.5 + 0.5) * 2048.)));
  return spread2((uint32_t)iu) | (spread2((uint32_t)iv) << 1);
}
CMIB_D uint32_t fine_position_key(const GridGeom &g, double px, double py, double pz) { /* 30 bits */
  const int ix = min(1023, max(0, (int)((px - g.anchor[0]) / g.sides[0] * 1024.)));
  const int iy = min(1023, max(0, (int)((py - g.anchor[1]) / g.sides[1] * 1024.)));
  const int iz = min(1023, max(0, (int)((pz - g.anchor[2]) / g.sides[2] * 1024.)));
  return spread3((uint32_t)ix) | (spread3((uint32_t)iy) << 1) | (spread3((uint32_t)iz) << 2);
}
/* bin counts -> bin offsets: exclusive prefix sum over nbins = ntiles * SORT_SCAN_TILE entries in three small
 * launches (tile sums, scan of the <= 2048 tile sums, tile-local scan + tile offset).  The last one also clears
 * the counts for the next round. */
constexpr int SORT_SCAN_BLOCK = 256;
constexpr int SORT_SCAN_ITEMS = 16;
constexpr int SORT_SCAN_TILE = SORT_SCAN_BLOCK * SORT_SCAN_ITEMS; /* 4096 bins per CTA */
constexpr int SORT_MAX_TILES = SORT_SCAN_BLOCK * 32;              /* what sort_scan_sums_kernel scans in one CTA */
constexpr int SORT_MAX_KEY_BITS = 23;                             /* + 1 flag bit: 2^24 bins, 64 MB: L2 resident (measured with
                                                                   * 2^25: the tickets of prepare_kernel thrash, +30 % prepare) */
static_assert((2ll << SORT_MAX_KEY_BITS) <= (long long)SORT_SCAN_TILE * SORT_MAX_TILES, "bin table larger than the scan");

CMIB_D uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
  __shared__ uint32_t warp_sums[SORT_SCAN_BLOCK / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += n;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    uint32_t s = (lane < SORT_SCAN_BLOCK / 32) ? warp_sums[lane] : 0u;
#pragma unroll
    for (int o = 1; o < SORT_SCAN_BLOCK / 32; o <<= 1) {
      const uint32_t n = __shfl_up_sync(0xffffffffu, s, o);
      if (lane >= o) s += n;
    }
    if (lane < SORT_SCAN_BLOCK / 32) warp_sums[lane] = s; /* inclusive */
  }
  __syncthreads();
  if (total) *total = warp_sums[SORT_SCAN_BLOCK / 32 - 1];
  return incl - v + (warp > 0 ? warp_sums[warp - 1] : 0u);
}
__global__ void __launch_bounds__(SORT_SCAN_BLOCK) sort_scan_tiles_kernel(const uint32_t *hist, uint32_t *block_sums) {
  const uint4 *h = reinterpret_cast<const uint4 *>(hist + (size_t)blockIdx.x * SORT_SCAN_TILE) + threadIdx.x * (SORT_SCAN_ITEMS / 4);
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SORT_SCAN_ITEMS / 4; ++k) { const uint4 v = h[k]; s += v.x + v.y + v.z + v.w; }
  uint32_t total;
  (void)block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}
/* one CTA: exclusive scan of the tile sums in place (ntiles <= SORT_MAX_TILES) */
__global__ void __launch_bounds__(SORT_SCAN_BLOCK) sort_scan_sums_kernel(uint32_t *block_sums, uint32_t ntiles) {
  constexpr int PER = SORT_MAX_TILES / SORT_SCAN_BLOCK;
  uint32_t v[PER], s = 0;
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t i = threadIdx.x * PER + k;
    v[k] = (i < ntiles) ? block_sums[i] : 0u;
    s += v[k];
  }
  uint32_t run = block_exclusive_scan(s, nullptr);
#pragma unroll
  for (int k = 0; k < PER; ++k) {
    const uint32_t i = threadIdx.x * PER + k;
    if (i < ntiles) block_sums[i] = run;
    run += v[k];
  }
}
__global__ void __launch_bounds__(SORT_SCAN_BLOCK)
sort_scan_offsets_kernel(uint32_t *hist, const uint32_t *block_sums, uint32_t *offs) {
  uint4 *h = reinterpret_cast<uint4 *>(hist + (size_t)blockIdx.x * SORT_SCAN_TILE) + threadIdx.x * (SORT_SCAN_ITEMS / 4);
  uint4 *o = reinterpret_cast<uint4 *>(offs + (size_t)blockIdx.x * SORT_SCAN_TILE) + threadIdx.x * (SORT_SCAN_ITEMS / 4);
  uint4 v[SORT_SCAN_ITEMS / 4];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < SORT_SCAN_ITEMS / 4; ++k) { v[k] = h[k]; s += v[k].x + v[k].y + v[k].z + v[k].w; }
  uint32_t run = block_sums[blockIdx.x] + block_exclusive_scan(s, nullptr);
#pragma unroll
  for (int k = 0; k < SORT_SCAN_ITEMS / 4; ++k) {
    uint4 r;
    r.x = run; run += v[k].x;
    r.y = run; run += v[k].y;
    r.z = run; run += v[k].z;
    r.w = run; run += v[k].w;
    o[k] = r;
    h[k] = make_uint4(0u, 0u, 0u, 0u);
  }
}
/* queue entry i goes to position offs[key] + its ticket */
__global__ void sort_scatter_kernel(const unsigned long long *ctl, const uint32_t *key, const uint32_t *rank,
                                    const uint32_t *offs, uint32_t *order) {
  const uint64_t n = ctl[CTL_QCOUNT];
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    order[offs[__ldcs(key + i)] + __ldcs(rank + i)] = (uint32_t)i;
}

/* meta word of a queue entry: uniforms consumed (32 bits) | packet type (7 bits) | emitted by the
 * continuous source (1 bit: the packet keeps that source's weight through re-emissions) | source
 * index + 1 of a primary of a discrete source, 0 otherwise (24 bits) */
constexpr int META_CONTINUOUS = 0x80; /* in the type byte */
CMIB_D uint64_t pack_meta(uint32_t ndraw, int type, int isrc = -1) {
  return ((uint64_t)(uint32_t)(isrc + 1) << 40) | ((uint64_t)(uint32_t)(type & 0xff) << 32) | ndraw;
}
CMIB_D int meta_type(uint64_t meta) { return (int)((meta >> 32) & 0x7fu); }
CMIB_D int meta_continuous(uint64_t meta) { return (int)((meta >> 32) & (uint64_t)META_CONTINUOUS); }
CMIB_D int meta_source(uint64_t meta) { return (int)(meta >> 40) - 1; }

constexpr int HOT_CELLS = 27;      /* 3 x 3 x 3 neighbourhood of a source cell */
constexpr int HOT_STRIDE = 16;     /* doubles per replicated cell record (one 128-B line) */
constexpr int HOT_CROSSINGS = 3;   /* crossings after emission that may still be in the neighbourhood */
constexpr int HOT_MAX_SOURCES = 64;

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

/* block-wide sum of per-thread counters into the 9 leading doubles of acc */
CMIB_D void reduce_counters(double *acc, const ShootCounters &cnt) {
  __shared__ double red[9][32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double v[9] = {cnt.w_tot, cnt.w_type[0], cnt.w_type[1], cnt.w_type[2], cnt.w_type[3],
                 (double)cnt.n_steps, (double)cnt.n_emit, (double)cnt.n_red, cnt.tau_sum};
#pragma unroll
  for (int k = 0; k < 9; ++k) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < 9) {
    double sum = 0.;
    for (int w = 0; w < (int)((blockDim.x + 31) >> 5); ++w) sum += red[threadIdx.x][w];
    if (sum != 0.) atomicAdd(acc + threadIdx.x, sum);
  }
}

/* ------------------------------------------------------------------------- */
/* re-emission decision: re-emission queue -> emission queue                   */
/* ------------------------------------------------------------------------- */
/* PhotonSource::reemit up to the choice of the new frequency (PhotonSource.cpp:272-295,
 * PhysicalDiffuseReemissionHandler.cpp:219-370).  Cheap and branchy; its survivors are compacted
 * into the emission queue so that the expensive, uniform part of an emission (direction, 14 cross
 * sections, optical depth) runs on full warps in prepare_kernel.  (With both in one kernel the
 * mixed rounds ran at 10 active threads per instruction, profiles/r01_prepare.md.) */
#ifndef CMIB_PREP_BLOCKS
#define CMIB_PREP_BLOCKS 4
#endif
#ifndef CMIB_DECIDE_BLOCKS
#define CMIB_DECIDE_BLOCKS 4 /* 64 registers, no spills: emission kernels of a lexingtonHII20 step 21.45 -> 20.77 ms against 3 CTAs (80 registers); 5 CTAs (48 registers, spills): 22.2 ms */
#endif
enum EmitField : int { EQ_PX = 0, EQ_PY, EQ_PZ, EQ_NU, EQ_ID, EQ_META, EQ_NFIELDS };

__global__ void __launch_bounds__(256, CMIB_DECIDE_BLOCKS)
reemit_decide_kernel(const __grid_constant__ WavefrontParams W) {
  const ShootParams &P = W.sp;
  const SourceModel &m = P.src;
  const uint64_t cap = W.capacity;
  const uint64_t n_re = W.ctl[CTL_RQCOUNT];
  ShootCounters cnt;
  const int lane = threadIdx.x & 31;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t first = (uint64_t)blockIdx.x * blockDim.x + (threadIdx.x & ~31u);
  for (uint64_t w0 = first; w0 < n_re; w0 += stride) {
    const uint64_t w = w0 + lane;
    bool emit = false;
    PacketRng rng;
    double px = 0., py = 0., pz = 0., nu = 0.;
    uint64_t id = 0;
    int type = PACKET_ABSORBED;
    int cont = 0;
    if (w < n_re) {
      px = __ldcs(W.rq + RQ_PX * cap + w);
      py = __ldcs(W.rq + RQ_PY * cap + w);
      pz = __ldcs(W.rq + RQ_PZ * cap + w);
      const double sigH = __ldcs(W.rq + RQ_SIGH * cap + w);
      const double sigHe = __ldcs(W.rq + RQ_SIGHE * cap + w);
      const int64_t cell = __double_as_longlong(__ldcs(W.rq + RQ_CELL * cap + w));
      id = (uint64_t)__double_as_longlong(__ldcs(W.rq + RQ_ID * cap + w));
      const uint64_t meta = (uint64_t)__double_as_longlong(__ldcs(W.rq + RQ_META * cap + w));
      rng_restore(rng, P.seed, P.iteration, id, (uint32_t)meta);
      type = meta_type(meta);
      cont = meta_continuous(meta);
      if (m.reemission_kind == REEMISSION_PHYSICAL) {
        const CellOpacity c = load_cell(P.cells, cell);
        /* the five cumulative probabilities are read where the decision needs them: an absorption by
         * hydrogen (most of them) looks at the first one only */
        nu = physical_reemit(m, sigH, sigHe, c.xH, c.xHe, c.T, P.reemit_prob + cell * NUM_REEMIT, rng, type);
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
        const double weight = cont ? m.continuous_weight : m.discrete_weight;
        cnt.w_tot += weight;
#pragma unroll
        for (int t = 0; t < NUM_PACKET_TYPES; ++t) cnt.w_type[t] += (t == type) ? weight : 0.;
      } else {
        emit = true;
      }
    }
    const unsigned ballot = __ballot_sync(0xffffffffu, emit);
    if (ballot) {
      unsigned long long base = 0;
      const int leader = __ffs(ballot) - 1;
      if (lane == leader) base = atomicAdd(&W.ctl[CTL_EQCOUNT], (unsigned long long)__popc(ballot));
      base = __shfl_sync(0xffffffffu, base, leader);
      if (emit) {
        double *q = W.eq + (base + __popc(ballot & ((1u << lane) - 1u)));
        __stcs(q + EQ_PX * cap, px); __stcs(q + EQ_PY * cap, py); __stcs(q + EQ_PZ * cap, pz);
        __stcs(q + EQ_NU * cap, nu);
        __stcs(q + EQ_ID * cap, __longlong_as_double((long long)id));
        __stcs(q + EQ_META * cap, __longlong_as_double((long long)pack_meta(rng_save(rng), type | cont)));
      }
    }
  }
  reduce_counters(P.acc, cnt);
}

/* ------------------------------------------------------------------------- */
/* tail: the last generations of re-emitted packets, one thread per packet     */
/* ------------------------------------------------------------------------- */
/* Once the primaries are out, every round holds ~0.36 x the packets of the round before (lexingtonHII20): from a few
 * 1e4 packets on a round costs its fixed price (five launches, one walk latency), ~25 of them per shoot — 1.5 ms of a
 * 15 ms shoot at 8 GPUs (profiles/r02_exchange.md).  When no primaries remain and the re-emission queue holds at most
 * TAIL_MAX packets, this kernel follows each of them to its end (packet_walks of shoot.cuh: the very code the
 * per-packet kernel runs, entered at the re-emission decision) and empties the queue; the round then finds nothing to
 * emit and the shoot ends.  Launched at the head of every round; a no-op until its condition holds. */
constexpr unsigned long long TAIL_MAX = 65536;
constexpr int TAIL_BLOCK = 128;
struct CountingAdder {
  uint32_t *n;
  __device__ __forceinline__ void operator()(double *addr, double v) const {
    atomicAdd(addr, v);
    ++*n;
  }
};
template <int MODE>
__global__ void __launch_bounds__(TAIL_BLOCK)
tail_kernel(const __grid_constant__ WavefrontParams W) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  const ShootParams &P = W.sp;
  const SourceModel &m = P.src;
  const uint64_t cap = W.capacity;
  const uint64_t n_re = W.ctl[CTL_RQCOUNT];
  if (n_re == 0 || n_re > TAIL_MAX || W.ctl[CTL_REMAINING] != 0) return; /* uniform over the grid */
  ShootCounters cnt;
  const CountingAdder add = {&cnt.n_red};
  const uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w < n_re) {
    MarchState s;
    s.px = __ldcs(W.rq + RQ_PX * cap + w);
    s.py = __ldcs(W.rq + RQ_PY * cap + w);
    s.pz = __ldcs(W.rq + RQ_PZ * cap + w);
    s.dx = s.dy = s.dz = 0.;
    double sigma[NSIG];
#pragma unroll
    for (int k = 0; k < NSIG; ++k) sigma[k] = 0.;
    sigma[0] = __ldcs(W.rq + RQ_SIGH * cap + w);
    if (NSIG > 1) sigma[(NSIG > 1) ? ION_He_n : 0] = __ldcs(W.rq + RQ_SIGHE * cap + w);
    double sigma_He_corr = 0.; /* set by the cross sections of the re-emitted packet before it is used */
    s.last_cell = __double_as_longlong(__ldcs(W.rq + RQ_CELL * cap + w));
    const uint64_t id = (uint64_t)__double_as_longlong(__ldcs(W.rq + RQ_ID * cap + w));
    const uint64_t meta = (uint64_t)__double_as_longlong(__ldcs(W.rq + RQ_META * cap + w));
    PacketRng rng;
    rng_restore(rng, P.seed, P.iteration, id, (uint32_t)meta);
    int type = meta_type(meta);
    const double weight = meta_continuous(meta) ? m.continuous_weight : m.discrete_weight;
    double nu = 0.;
    const CellOpacity c = load_cell(P.cells, s.last_cell);
    packet_walks<MODE>(P, rng, add, cnt, s, sigma, sigma_He_corr, nu, type, weight, true, c);
  }
  reduce_counters(P.acc, cnt);
  /* the CTA that finishes last empties the queue (every CTA has read its size by then) */
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned long long done = atomicAdd(&W.ctl[CTL_TAIL_BLOCKS], 1ull);
    if (done == (unsigned long long)gridDim.x - 1ull) {
      W.ctl[CTL_TAIL_BLOCKS] = 0ull;
      W.ctl[CTL_RQCOUNT] = 0ull;
    }
  }
}

/* ------------------------------------------------------------------------- */
/* prepare: emission queue + fresh primaries -> march queue                    */
/* ------------------------------------------------------------------------- */
/* Every item becomes a packet: item w goes to march-queue slot w (no atomics).  Items
 * [0, n_eq) are re-emissions (frequency already drawn), [n_eq, n_eq + n_fresh) are primaries
 * (PhotonSource::get_random_photon, PhotonSource.cpp:208-249); both then draw the direction
 * (PhotonSource.hpp:141-148), evaluate the 14 cross sections (set_cross_sections, :189-199) and the
 * optical depth tau = -ln u (IonizationPhotonShootJob.hpp:135). */
template <int MODE>
__global__ void __launch_bounds__(256, CMIB_PREP_BLOCKS) /* 4 CTAs per SM (64 registers, 36 bytes of spills).  Without a bound
                                                          * the rarely taken continuous-source branches of emit_primary push
                                                          * it to 104 registers (2 CTAs, +2 ms per lexingtonHII20 step); 3
                                                          * CTAs (80 registers): emission kernels 20.77 ms per step, 4: 20.3 */
prepare_kernel(const __grid_constant__ WavefrontParams W) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  const ShootParams &P = W.sp;
  const SourceModel &m = P.src;
  const uint64_t cap = W.capacity;
  const uint64_t n_eq = W.ctl[CTL_EQCOUNT];
  const uint64_t remaining = W.ctl[CTL_REMAINING];
  const uint64_t next_fresh = W.ctl[CTL_NEXT_FRESH];
  const uint64_t room = W.fill > n_eq ? W.fill - n_eq : 0;
  const uint64_t n_fresh = remaining < room ? remaining : room;
  const uint64_t n_items = n_eq + n_fresh;
  ShootCounters cnt;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w < n_items; w += stride) {
    PacketRng rng;
    double px, py, pz, nu, dx, dy, dz;
    uint64_t id;
    int type = PACKET_PRIMARY; /* | META_CONTINUOUS */
    int isrc_key = -1; /* source index of a primary of a discrete source, -1 otherwise */
    if (w < n_eq) {
      px = __ldcs(W.eq + EQ_PX * cap + w);
      py = __ldcs(W.eq + EQ_PY * cap + w);
      pz = __ldcs(W.eq + EQ_PZ * cap + w);
      nu = __ldcs(W.eq + EQ_NU * cap + w);
      id = (uint64_t)__double_as_longlong(__ldcs(W.eq + EQ_ID * cap + w));
      const uint64_t meta = (uint64_t)__double_as_longlong(__ldcs(W.eq + EQ_META * cap + w));
      rng_restore(rng, P.seed, P.iteration, id, (uint32_t)meta);
      type = meta_type(meta) | meta_continuous(meta);
      random_direction(rng, dx, dy, dz);
    } else {
      id = P.packet_offset + next_fresh + (w - n_eq);
      rng_init(rng, P.seed, P.iteration, id);
      emit_primary(m, P.geom, rng, px, py, pz, dx, dy, dz, nu, isrc_key);
      if (isrc_key < 0) type |= META_CONTINUOUS;
    }
    double sigma_He_corr;
    double sigma[NSIG];
    ++cnt.n_emit;
    packet_cross_sections<NSIG>(m, nu, sigma, sigma_He_corr);
    const double u_tau = rng_uniform(rng);
    const double tau = -log(u_tau);
    /* streaming stores: a queue entry is read once, by the march, after the rest of the round has been written */
    double *q = W.mq + w;
    __stcs(q + MQ_PX * cap, px); __stcs(q + MQ_PY * cap, py); __stcs(q + MQ_PZ * cap, pz);
    __stcs(q + MQ_DX * cap, dx); __stcs(q + MQ_DY * cap, dy); __stcs(q + MQ_DZ * cap, dz);
    __stcs(q + MQ_NU * cap, nu); __stcs(q + MQ_TAU * cap, tau);
    __stcs(q + MQ_ID * cap, __longlong_as_double((long long)id));
    __stcs(q + MQ_META * cap, __longlong_as_double((long long)pack_meta(rng_save(rng), type, isrc_key)));
#pragma unroll
    for (int k = 0; k < NSIG; ++k) __stcs(q + (MQ_SIGMA + k) * cap, sigma[k]);
    if (MODE == ACC_FULL) __stcs(q + (MQ_SIGMA + NSIG) * cap, sigma_He_corr);
    if (W.sort == 2) {
      /* [re-emitted] [source | direction, or position] [optical-depth bin]: neighbours in key order leave in nearly
       * the same direction AND carry nearly the same optical depth, i.e. they are absorbed near the same place: the
       * 8 lanes of a refill group finish together instead of waiting for the deepest of 8 random depths */
      const int kb = W.fine_key_bits - W.tau_bits;
      uint32_t k = (isrc_key >= 0)
                       ? (((uint32_t)isrc_key << W.fine_dir_bits) | (fine_direction_key(dx, dy, dz) >> (22 - W.fine_dir_bits)))
                       : ((1u << kb) | (fine_position_key(P.geom, px, py, pz) >> (30 - kb)));
      k = (k << W.tau_bits) | (uint32_t)min((int)(u_tau * (double)(1 << W.tau_bits)), (1 << W.tau_bits) - 1);
      __stcs(W.key + w, k);
      __stcs(W.rank + w, atomicAdd(&W.hist[k], 1u)); /* a ticket inside the bin: counting sort without a second pass */
    }
  }
  reduce_counters(P.acc, cnt);
}

/* bookkeeping after prepare: consume the primaries, empty the re-emission queue */
__global__ void advance_after_prepare_kernel(unsigned long long *ctl, uint64_t fill) {
  const uint64_t n_eq = ctl[CTL_EQCOUNT];
  const uint64_t remaining = ctl[CTL_REMAINING];
  const uint64_t room = fill > n_eq ? fill - n_eq : 0;
  const uint64_t n_fresh = remaining < room ? remaining : room;
  ctl[CTL_REMAINING] = remaining - n_fresh;
  ctl[CTL_NEXT_FRESH] += n_fresh;
  ctl[CTL_QCOUNT] = n_eq + n_fresh; /* every item of prepare became a packet */
  ctl[CTL_RQCOUNT] = 0;
  ctl[CTL_EQCOUNT] = 0;
  ctl[CTL_HEAD] = 0;
  const uint64_t round = ctl[CTL_ROUND];
  ctl[CTL_STATUS + (round % CTL_STATUS_SLOTS)] = ctl[CTL_QCOUNT];
  ctl[CTL_ROUND] = round + 1;
}

__global__ void advance_after_march_kernel(unsigned long long *ctl, const double *acc) {
  ctl[CTL_QCOUNT] = 0;
  ctl[CTL_CROSSINGS] = (unsigned long long)__double_as_longlong(acc[5]);
  ctl[CTL_EMISSIONS] = (unsigned long long)__double_as_longlong(acc[6]);
}
__global__ void shoot_begin_kernel(unsigned long long *ctl, const double *acc) {
  ctl[CTL_CROSSINGS0] = ctl[CTL_CROSSINGS] = (unsigned long long)__double_as_longlong(acc[5]);
  ctl[CTL_EMISSIONS0] = ctl[CTL_EMISSIONS] = (unsigned long long)__double_as_longlong(acc[6]);
}

/* fold the hot-cell replicas into the accumulators and clear them (one thread per
 * (source, neighbour cell, term); the replicas of one record are summed in a fixed order) */
template <int MODE>
__global__ void fold_hot_cells_kernel(const __grid_constant__ ShootParams P) {
  const int n = P.src.n_sources * HOT_CELLS * HOT_STRIDE;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int k = t % HOT_STRIDE, nb = (t / HOT_STRIDE) % HOT_CELLS, isrc = t / (HOT_STRIDE * HOT_CELLS);
  double sum = 0.;
  for (int r = 0; r < P.hot_replicas; ++r) {
    double *p = P.hot_acc + (((size_t)r * HOT_MAX_SOURCES + isrc) * HOT_CELLS + nb) * HOT_STRIDE + k;
    sum += *p;
    *p = 0.;
  }
  if (sum == 0.) return;
  const uint32_t sc = P.src_cell[isrc];
  const int ix = (int)(sc & 1023u) + nb / 9 - 1, iy = (int)((sc >> 10) & 1023u) + (nb / 3) % 3 - 1,
            iz = (int)((sc >> 20) & 1023u) + nb % 3 - 1;
  /* a neighbour outside the grid never received anything (the walk only visits cells inside) */
  const int64_t cell = long_index(P.geom, ix, iy, iz);
  /* full layout: a replica record has the layout of a cell record (k is a slot); H-only: k is the
   * term, 0 (J_H) or 1 (heat_H) */
  if (MODE == ACC_FULL) atomicAdd(P.acc + ACC_COUNTERS + cell * AccLayout<MODE>::NACC + k, sum);
  else atomicAdd(acc_term<MODE>(P, cell, k), sum);
}

/* ------------------------------------------------------------------------- */
/* march: persistent warp state machine over the march queue                  */
/* ------------------------------------------------------------------------- */
constexpr int MARCH_BLOCK = 256;
constexpr int MARCH_CHUNK = 128;     /* queue entries a warp claims with one atomic */
constexpr int MARCH_REFILL_MIN = 8;  /* lanes waiting (walk ended or empty) that trigger finish + refill */

enum LaneState : int { LANE_EMPTY = 0, LANE_LIVE = 1, LANE_ABSORBED = 2, LANE_ESCAPED = 3 };

/*
 * The hot loop executes ONE cell crossing per pass for every live lane and nothing else.
 * Everything that happens once per packet — the partial-step correction with its four
 * divisions (CartesianDensityGrid.cpp:413-417), the accumulation of that last partial
 * step, the hand-over to the re-emission queue, loading the next packet — is deferred
 * until at least MARCH_REFILL_MIN lanes wait for it, so that it runs on many lanes at
 * once instead of on ~1 (ncu: profiles/r01_march_v1.md).
 *
 * Arithmetic of the crossing is that of march_step (march.cuh), operation for operation;
 * the only re-arrangements are exact: (double)ix is carried as a double that is
 * incremented by +-1 (integers are exact in FP64), and the upper wall lo + cellside of
 * the reference is formed as lo + h with h = cellside (d > 0) or 0 (d < 0), x + 0 == x.
 */
/*
 * AGG (coherent march, W.sort == 2): the queue is ordered so that the lanes of a warp walk
 * through the same cells; lanes that add to the same accumulator record are found with
 * match.any, their J_H / heat_H (/ J_O0 / J_N0: the first sector of a full record) terms are
 * summed in registers (reduce_peers) and the lowest lane issues ONE RED per term.  The gathers of
 * such lanes fall into one sector by themselves.  The remaining terms (photons above 21.6 eV)
 * are added per lane as in the plain kernel.
 */
constexpr uint32_t AGG_METALS = (1u << ION_O_n) | (1u << ION_N_n);
constexpr uint32_t MASK_METALS = 0x3ffcu;        /* bits 2..13 of a lane's mask word */
constexpr uint32_t MASK_CONTINUOUS = 1u << 31;   /* packet of the continuous source: its weight differs */
constexpr int AGG_GROUP = 8; /* lanes that are refilled together */
/*
 * PRE: the cell record of the NEXT crossing is requested as soon as the packet's next cell is
 * known (end of a crossing / end of a refill) and consumed one pass later, so that the gather's
 * L2/DRAM latency overlaps the in-warp sums, the loop control and the wall distances of the next
 * pass instead of stalling the optical-depth arithmetic right behind the load.
 */
#ifndef CMIB_PLAIN_BLOCKS
#define CMIB_PLAIN_BLOCKS 3 /* resident CTAs per SM of the plain kernel (80 registers).  The walk is sensitive to both ends:
                             * 2 CTAs (grid halved): lexingtonHII20 march 74.9 -> 95.9 ms; 4 CTAs (64 registers, 130 bytes of
                             * spills in the full layout): 94.4 ms */
#endif
#ifndef CMIB_AGG_BLOCKS
#define CMIB_AGG_BLOCKS 3 /* resident CTAs per SM the coherent variants are compiled for.  Measured with 2 (118
                           * registers, no spills): H-only 26.9 -> 32.2 ms on clumpy 256^3, full layout 117 -> 114 ms:
                           * the 24 warps per SM matter more than the spills */
#endif
template <int MODE, bool AGG, bool PRE>
__global__ void __launch_bounds__(MARCH_BLOCK, (AGG ? CMIB_AGG_BLOCKS : CMIB_PLAIN_BLOCKS))
march_kernel(const __grid_constant__ WavefrontParams W) {
  constexpr int NSIG = AccLayout<MODE>::NSIG;
  constexpr int NMETAL = (MODE == ACC_FULL) ? 12 : 0;
  constexpr int NV = (MODE == ACC_FULL) ? 4 : 2; /* terms summed inside the warp (AGG) */
  /* per-lane packet constants that are only touched once per packet or per accumulation */
  __shared__ double s_sig[(NMETAL > 0 ? NMETAL + 1 : 1)][MARCH_BLOCK]; /* metals, then sigma_He */
  __shared__ unsigned long long s_id[MARCH_BLOCK], s_meta[MARCH_BLOCK];
  __shared__ double s_tau0[MARCH_BLOCK]; /* sampled optical depth, for the traversed-depth checksum */
  __shared__ uint32_t s_ntype_c[NUM_PACKET_TYPES][MARCH_BLOCK]; /* ended packets of the continuous source */
  const ShootParams &P = W.sp;
  const GridGeom &g = P.geom;
  const uint64_t cap = W.capacity;
  const uint64_t qcount = W.ctl[CTL_QCOUNT];
  const int lane = threadIdx.x & 31;
  const int tid = threadIdx.x;
  const bool can_reemit = (P.src.reemission_kind != REEMISSION_NONE);
  const bool any_periodic = (g.periodic[0] | g.periodic[1] | g.periodic[2]) != 0;
  const double w_discrete = P.src.discrete_weight, w_continuous = P.src.continuous_weight;
#pragma unroll
  for (int t = 0; t < NUM_PACKET_TYPES; ++t) s_ntype_c[t][tid] = 0u;
  const uint32_t ncx = (uint32_t)g.ncell[0], ncy = (uint32_t)g.ncell[1], ncz = (uint32_t)g.ncell[2];
  uint32_t n_type[NUM_PACKET_TYPES] = {0u, 0u, 0u, 0u};
  uint32_t n_steps = 0;
  uint32_t mask = 0; /* metals (bits 2..13) with a non-zero cross section | MASK_CONTINUOUS */
  /* a packet ended with type t: packets of the two kinds of source carry different weights */
  auto count_end = [&](int t) {
    if (mask & MASK_CONTINUOUS) ++s_ntype_c[t][tid];
    else ++n_type[t];
  };

  /* packet state */
  double px = 0., py = 0., pz = 0., dx = 1., dy = 1., dz = 1., ivx = 1., ivy = 1., ivz = 1.;
  double fx = 0., fy = 0., fz = 0.; /* cell indices as doubles */
  double tau = 0., tau_cell = 0., ds = 0.;
  double sigH = 0., sigHe_corr = 0., dnu_H = 0., dnu_He = 0.;
  int32_t ix = 0, iy = 0, iz = 0;
  uint32_t cell = 0;
  uint32_t hot = 0;  /* (hot record index of the packet's source + 1) << 2 | crossings made (saturating) */
  uint32_t hot_cell = 0; /* packed cell indices of that source */
  uint32_t nacc = 0; /* accumulator terms this packet adds per crossing (diagnostic for the roofline) */
  uint32_t n_red = 0;
  double tau_sum = 0.; /* optical depth traversed (checksum against sum_cells n (x_H J_H + A_He x_He J_He)) */
  int state = LANE_EMPTY;
  /* PRE: record of the cell the lane is in.  (ptxas waits for the outstanding request at the loop-head
   * branch, 12 % of the stall samples, because the refill path writes these registers too; loading the
   * first cell of a packet on the crossing path instead removed that wait but cost a branch and spills:
   * clumpy 256^3 26.3 -> 26.9 ms, so it stays as it is.) */
  double pre_n = 0., pre_xH = 0., pre_xHe = 0.;
  auto prefetch_cell = [&]() {
    const uint32_t pc = ((uint32_t)ix * ncy + (uint32_t)iy) * ncz + (uint32_t)iz;
    if (MODE == ACC_HONLY) {
      const double2 r0 = __ldg(P.cells_h + pc);
      pre_n = r0.x; pre_xH = r0.y;
    } else {
      double t;
      asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
          : "=d"(pre_n), "=d"(pre_xH), "=d"(pre_xHe), "=d"(t)
          : "l"(P.cells + pc));
    }
  };
  bool warp_has_zero_dir = false; /* some lane's direction has a zero component (warp-uniform) */
  uint64_t cur = 0, end = 0;      /* warp-uniform cursor into the claimed chunk */
  const uint64_t nchunks = (qcount + MARCH_CHUNK - 1) / MARCH_CHUNK;
  bool exhausted = (qcount == 0);
  uint32_t n_pass = 0;

  while (true) {
    /* safety valve: a warp needs ~(entries per warp) x (crossings per packet) passes, orders
     * of magnitude below this bound; never spin forever on a shared GPU */
    if (++n_pass > (1u << 26)) {
      if (lane == 0) atomicExch(&W.ctl[CTL_ERROR], 1ull);
      break;
    }
    const unsigned live_m = __ballot_sync(0xffffffffu, state == LANE_LIVE);
    const unsigned waiting = ~live_m;               /* empty lanes + lanes whose walk ended */
    const int nwait = __popc(waiting);
    /* service (finish + refill) when enough lanes wait for it or when nothing can be stepped;
     * once the queue is exhausted only lanes with a pending finish count */
    /* AGG: lanes are refilled in aligned groups of AGG_GROUP lanes, a group when all of its
     * lanes wait, so that lane order inside a group stays queue (= key) order */
    unsigned fill_m = waiting;
    if (AGG) {
      static_assert(AGG_GROUP == 8, "group mask arithmetic below is written for 8 lanes");
      unsigned m = waiting;
      m &= m >> 1;
      m &= m >> 2;
      m &= m >> 4;
      fill_m = (m & 0x01010101u) * 0xffu;
    }
    bool service = AGG ? (fill_m != 0u) : (nwait >= MARCH_REFILL_MIN);
    if (service && exhausted) {
      const unsigned pend_m = __ballot_sync(0xffffffffu, state == LANE_ABSORBED || state == LANE_ESCAPED);
      if (live_m == 0u && pend_m == 0u) break;
      service = (pend_m != 0u) && (__popc(pend_m) >= MARCH_REFILL_MIN || live_m == 0u);
    }
    if (service) {
      /* ---- finish: absorbed ---- */
      double fpx = 0., fpy = 0., fpz = 0.;
      if (state == LANE_ABSORBED && !(tau < 0.)) {
        /* tau == 0 exactly after a full crossing: the walk ends on the wall, in the cell the
         * packet has just entered (interact() returns the cell of the current index, :445-451) */
        fpx = px; fpy = py; fpz = pz;
        cell = ((uint32_t)ix * ncy + (uint32_t)iy) * ncz + (uint32_t)iz;
        if (!can_reemit) count_end(PACKET_ABSORBED);
      } else if (state == LANE_ABSORBED) {
        /* tau < 0 after the crossing: shorten it (CartesianDensityGrid.cpp:413-417) */
        const double nwx = xadd(px, xmul(ds, dx));
        const double nwy = xadd(py, xmul(ds, dy));
        const double nwz = xadd(pz, xmul(ds, dz));
        const double Scorr = xdiv(xmul(ds, tau), tau_cell);
        const double dss = xadd(ds, Scorr);
        fpx = xadd(px, xdiv(xmul(xsub(nwx, px), dss), ds));
        fpy = xadd(py, xdiv(xmul(xsub(nwy, py), dss), ds));
        fpz = xadd(pz, xdiv(xmul(xsub(nwz, pz), dss), ds));
        /* accumulate the shortened crossing; the cell has n > 0 (tau_cell > 0) */
        n_red += nacc;
        if (AGG) n_red += (sigH != 0.) * (1u + (dnu_H != 0.)) + __popc(mask & AGG_METALS); /* added per lane here */
        const double dsw = dss * ((mask & MASK_CONTINUOUS) ? w_continuous : w_discrete);
        double *a = P.acc + ACC_COUNTERS + (size_t)cell * AccLayout<MODE>::NACC;
        const double dJH = dsw * sigH;
        if (dJH != 0.) {
          atomicAdd(acc_term<MODE>(P, cell, ION_H_n), dJH);
          const double dh = dJH * dnu_H;
          if (dh != 0.) atomicAdd(acc_term<MODE>(P, cell, MODE == ACC_FULL ? NUM_IONS + HEAT_H : 1), dh);
        }
        if (MODE == ACC_FULL) {
          const double dJHe = dsw * s_sig[NMETAL][tid];
          if (dJHe != 0.) {
            atomicAdd(a + acc_slot(ION_He_n), dJHe);
            const double dh = dJHe * dnu_He;
            if (dh != 0.) atomicAdd(a + acc_slot(NUM_IONS + HEAT_He), dh);
          }
          uint32_t mm = mask & MASK_METALS;
          while (mm) {
            const int k = __ffs(mm) - 1;
            mm &= mm - 1u;
            const double dJ = dsw * s_sig[k - 2][tid];
            if (dJ != 0.) atomicAdd(a + acc_slot(k), dJ);
          }
        }
        if (!can_reemit) count_end(PACKET_ABSORBED); /* PhotonSource::reemit without a handler (:304-306) */
      } else if (state == LANE_ESCAPED) {
        count_end(meta_type(s_meta[tid])); /* keeps its last type (IonizationPhotonShootJob.hpp:143-144) */
      }
      /* optical depth traversed by the walk that just ended: all of it when absorbed */
      if (state == LANE_ABSORBED || state == LANE_ESCAPED) tau_sum += s_tau0[tid] - ((tau > 0.) ? tau : 0.);
      if (can_reemit) {
        const unsigned ab = __ballot_sync(0xffffffffu, state == LANE_ABSORBED);
        if (ab) {
          unsigned long long base = 0;
          const int leader = __ffs(ab) - 1;
          if (lane == leader) base = atomicAdd(&W.ctl[CTL_RQCOUNT], (unsigned long long)__popc(ab));
          base = __shfl_sync(0xffffffffu, base, leader);
          if (state == LANE_ABSORBED) {
            double *q = W.rq + (base + __popc(ab & ((1u << lane) - 1u)));
            __stcs(q + RQ_PX * cap, fpx); __stcs(q + RQ_PY * cap, fpy); __stcs(q + RQ_PZ * cap, fpz);
            __stcs(q + RQ_SIGH * cap, sigH);
            __stcs(q + RQ_SIGHE * cap, (MODE == ACC_FULL) ? s_sig[NMETAL][tid] : 0.);
            __stcs(q + RQ_CELL * cap, __longlong_as_double((long long)cell));
            __stcs(q + RQ_ID * cap, __longlong_as_double((long long)s_id[tid]));
            __stcs(q + RQ_META * cap, __longlong_as_double((long long)(s_meta[tid] & 0xffffffffffull)));
          }
        }
      }
      if (state != LANE_LIVE) state = LANE_EMPTY;

      /* ---- refill: hand queue entries to the empty lanes ---- */
      if (!exhausted) {
        if (cur == end) {
          unsigned long long b = 0;
          if (lane == 0) b = atomicAdd(&W.ctl[CTL_HEAD], (unsigned long long)MARCH_CHUNK);
          b = __shfl_sync(0xffffffffu, b, 0);
          if (b >= qcount) {
            exhausted = true;
          } else {
            /* experiment switch (CMIB_CHUNK_STRIDE, default 1 = in order): consecutive chunks of the
             * ordered queue walk the same cone; a stride coprime to nchunks spreads the warps in
             * flight over different cones.  Measured: in order is faster (shared L2 lines). */
            if (W.chunk_stride > 1u) b = (((b / MARCH_CHUNK) * (uint64_t)W.chunk_stride) % nchunks) * MARCH_CHUNK;
            cur = b;
            end = (b + MARCH_CHUNK < qcount) ? b + MARCH_CHUNK : qcount;
          }
        }
        if (!exhausted) {
          const int rank = __popc(fill_m & ((1u << lane) - 1u));
          const uint64_t avail = end - cur;
          const int nfill = __popc(fill_m);
          if (state == LANE_EMPTY && ((fill_m >> lane) & 1u) && (uint64_t)rank < avail) {
            const uint64_t slot = W.sort ? (uint64_t)W.order[cur + rank] : (cur + rank);
            const double *q = W.mq + slot;
            /* streaming loads: a queue entry is read once */
            px = __ldcs(q + MQ_PX * cap); py = __ldcs(q + MQ_PY * cap); pz = __ldcs(q + MQ_PZ * cap);
            dx = __ldcs(q + MQ_DX * cap); dy = __ldcs(q + MQ_DY * cap); dz = __ldcs(q + MQ_DZ * cap);
            const double nu = __ldcs(q + MQ_NU * cap);
            tau = __ldcs(q + MQ_TAU * cap);
            s_tau0[tid] = tau;
            s_id[tid] = (unsigned long long)__double_as_longlong(__ldcs(q + MQ_ID * cap));
            s_meta[tid] = (unsigned long long)__double_as_longlong(__ldcs(q + MQ_META * cap));
            hot = 0;
            if (P.hot_replicas > 0) {
              const int isrc = meta_source(s_meta[tid]);
              if (isrc >= 0) {
                hot = (uint32_t)(isrc + 1) << 2;
                hot_cell = P.src_cell[isrc];
              }
            }
            sigH = __ldcs(q + MQ_SIGMA * cap);
            mask = meta_continuous(s_meta[tid]) ? MASK_CONTINUOUS : 0u;
            if (MODE == ACC_FULL) {
              s_sig[NMETAL][tid] = __ldcs(q + (MQ_SIGMA + 1) * cap);
              sigHe_corr = __ldcs(q + (MQ_SIGMA + NSIG) * cap);
#pragma unroll
              for (int k = 0; k < NMETAL; ++k) {
                const double v = __ldcs(q + (MQ_SIGMA + 2 + k) * cap);
                s_sig[k][tid] = v;
                mask |= (v != 0.) ? (1u << (2 + k)) : 0u;
              }
            }
            dnu_H = nu - P.nu_H;
            dnu_He = nu - P.nu_He;
            /* AGG: the warp-summed terms are counted where they are issued */
            nacc = AGG ? __popc(mask & MASK_METALS & ~AGG_METALS) : (sigH != 0.) * (1u + (dnu_H != 0.)) + __popc(mask & MASK_METALS);
            if (MODE == ACC_FULL) nacc += (s_sig[NMETAL][tid] != 0.) * (1u + (dnu_He != 0.));
            ivx = 1. / dx;
            ivy = 1. / dy;
            ivz = 1. / dz;
            /* get_cell_indices (CartesianDensityGrid.cpp:152-161) */
            ix = trunc_index(xmul(xsub(px, g.anchor[0]), g.inv_cellside[0]));
            iy = trunc_index(xmul(xsub(py, g.anchor[1]), g.inv_cellside[1]));
            iz = trunc_index(xmul(xsub(pz, g.anchor[2]), g.inv_cellside[2]));
            state = LANE_LIVE;
            if (any_periodic) {
              MarchState ms;
              ms.px = px; ms.py = py; ms.pz = pz; ms.ix = ix; ms.iy = iy; ms.iz = iz;
              const bool in = march_inside(g, ms);
              px = ms.px; py = ms.py; pz = ms.pz; ix = ms.ix; iy = ms.iy; iz = ms.iz;
              if (!in) state = LANE_ESCAPED;
            } else if ((uint32_t)ix >= ncx || (uint32_t)iy >= ncy || (uint32_t)iz >= ncz) {
              state = LANE_ESCAPED; /* emitted outside the box: interact() returns end() */
            }
            fx = (double)ix; fy = (double)iy; fz = (double)iz;
            if (PRE && state == LANE_LIVE) prefetch_cell();
          }
          cur += ((uint64_t)nfill < avail) ? (uint64_t)nfill : avail;
        }
      }
      warp_has_zero_dir = __any_sync(0xffffffffu, state == LANE_LIVE && (dx == 0. || dy == 0. || dz == 0.));
      continue;
    }

    /* ---- one cell crossing for every live lane ---- */
    bool do_acc = false;    /* AGG: this lane holds terms for the warp-level sum */
    double *arec = nullptr; /* its accumulator record */
    uint32_t akey = 0;      /* ... and a 32-bit name for it: the cell, or 2^31 | hot replica record */
    int64_t ats = 1;
    double v[NV];
#pragma unroll
    for (int k = 0; k < NV; ++k) v[k] = 0.;
    if (state == LANE_LIVE) {
      cell = ((uint32_t)ix * ncy + (uint32_t)iy) * ncz + (uint32_t)iz;
      /* one gather per crossing: the 32-byte cell record is one sector; the H-only walk needs
       * (n, x_H) only = one 16-byte load, the full walk takes the sector as one 256-bit load */
      CellOpacity c;
      if (PRE) {
        c.n = pre_n; c.xH = pre_xH; c.xHe = (MODE == ACC_HONLY) ? 0. : pre_xHe; c.T = 0.;
      } else if (MODE == ACC_HONLY) {
        const double2 r0 = __ldg(P.cells_h + cell);
        c.n = r0.x; c.xH = r0.y; c.xHe = 0.; c.T = 0.;
      } else {
        asm("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
            : "=d"(c.n), "=d"(c.xH), "=d"(c.xHe), "=d"(c.T)
            : "l"(P.cells + cell));
      }
      /* get_cell / get_wall_intersection (CartesianDensityGrid.cpp:170-176, 280-318) */
      const double lox = xadd(g.anchor[0], xmul(g.cellside[0], fx));
      const double loy = xadd(g.anchor[1], xmul(g.cellside[1], fy));
      const double loz = xadd(g.anchor[2], xmul(g.cellside[2], fz));
      const bool posx = dx > 0., posy = dy > 0., posz = dz > 0.;
      double wx = xmul(xsub(xadd(lox, posx ? g.cellside[0] : 0.), px), ivx);
      double wy = xmul(xsub(xadd(loy, posy ? g.cellside[1] : 0.), py), ivy);
      double wz = xmul(xsub(xadd(loz, posz ? g.cellside[2] : 0.), pz), ivz);
      if (warp_has_zero_dir) {
        if (dx == 0.) wx = DBL_MAX;
        if (dy == 0.) wy = DBL_MAX;
        if (dz == 0.) wz = DBL_MAX;
      }
      const double myz = (wz < wy) ? wz : wy;
      ds = (myz < wx) ? myz : wx;
      /* ds * n * (sigma_H*x_H + sigma_Hecorr*x_He), left to right (DensityGrid.hpp:129-133) */
      tau_cell = xmul(xmul(ds, c.n), xadd(xmul(sigH, c.xH), xmul(sigHe_corr, c.xHe)));
      tau = xsub(tau, tau_cell);
      ++n_steps;
      if (tau < 0.) {
        state = LANE_ABSORBED; /* position, ds, tau, tau_cell stay as they are for the deferred finish */
      } else {
        if (c.n > 0.) {
          /* update_integrals (DensityGrid.hpp:150-197); zero increments are skipped (exact) */
          n_red += nacc;
          const double dsw = ds * ((mask & MASK_CONTINUOUS) ? w_continuous : w_discrete);
          /* accumulator record of this cell: its own, or — during the first crossings of a
           * primary, inside the 3x3x3 cells around its source — one of the replicas */
          double *a = (MODE == ACC_FULL) ? P.acc + ACC_COUNTERS + (size_t)cell * AccLayout<MODE>::NACC
                                         : acc_term<MODE>(P, cell, 0);
          int64_t ts = (MODE == ACC_FULL) ? 1 : P.honly_term_stride; /* stride between terms */
          akey = cell;
          if (hot != 0u && (hot & 3u) < (uint32_t)HOT_CROSSINGS) {
            const int ddx = ix - (int)(hot_cell & 1023u), ddy = iy - (int)((hot_cell >> 10) & 1023u),
                      ddz = iz - (int)((hot_cell >> 20) & 1023u);
            if ((unsigned)(ddx + 1) < 3u && (unsigned)(ddy + 1) < 3u && (unsigned)(ddz + 1) < 3u) {
              const size_t rec = ((size_t)(blockIdx.x % P.hot_replicas) * HOT_MAX_SOURCES + ((hot >> 2) - 1u)) * HOT_CELLS +
                                 (size_t)((ddx + 1) * 9 + (ddy + 1) * 3 + (ddz + 1));
              a = P.hot_acc + rec * HOT_STRIDE;
              ts = 1;
              akey = 0x80000000u | (uint32_t)rec;
            }
            ++hot;
          }
          const double dJH = dsw * sigH;
          if constexpr (AGG) {
            do_acc = true;
            arec = a;
            ats = ts;
            v[0] = dJH;
            v[1] = dJH * dnu_H;
            if constexpr (MODE == ACC_FULL) {
              v[2] = (mask & (1u << ION_O_n)) ? dsw * s_sig[ION_O_n - 2][tid] : 0.;
              v[3] = (mask & (1u << ION_N_n)) ? dsw * s_sig[ION_N_n - 2][tid] : 0.;
            }
          } else if (dJH != 0.) {
            atomicAdd(a + (MODE == ACC_FULL ? acc_slot(ION_H_n) : 0), dJH);
            const double dh = dJH * dnu_H;
            if (dh != 0.) atomicAdd(a + (MODE == ACC_FULL ? (int64_t)acc_slot(NUM_IONS + HEAT_H) : ts), dh);
          }
          if (MODE == ACC_FULL) {
            const double dJHe = dsw * s_sig[NMETAL][tid];
            if (dJHe != 0.) {
              atomicAdd(a + acc_slot(ION_He_n), dJHe);
              const double dh = dJHe * dnu_He;
              if (dh != 0.) atomicAdd(a + acc_slot(NUM_IONS + HEAT_He), dh);
            }
            uint32_t mm = AGG ? (mask & MASK_METALS & ~AGG_METALS) : (mask & MASK_METALS);
            while (mm) {
              const int k = __ffs(mm) - 1;
              mm &= mm - 1u;
              const double dJ = dsw * s_sig[k - 2][tid];
              if (dJ != 0.) atomicAdd(a + acc_slot(k), dJ);
            }
          }
        }
        /* move to the wall, step the indices of every axis whose wall was hit */
        const bool hitx = (wx == ds), hity = (wy == ds), hitz = (wz == ds);
        px = xadd(px, xmul(ds, dx));
        py = xadd(py, xmul(ds, dy));
        pz = xadd(pz, xmul(ds, dz));
        ix += hitx ? (posx ? 1 : -1) : 0;
        iy += hity ? (posy ? 1 : -1) : 0;
        iz += hitz ? (posz ? 1 : -1) : 0;
        fx += hitx ? (posx ? 1. : -1.) : 0.;
        fy += hity ? (posy ? 1. : -1.) : 0.;
        fz += hitz ? (posz ? 1. : -1.) : 0.;
        if (any_periodic) {
          MarchState ms;
          ms.px = px; ms.py = py; ms.pz = pz; ms.ix = ix; ms.iy = iy; ms.iz = iz;
          const bool in = march_inside(g, ms);
          px = ms.px; py = ms.py; pz = ms.pz;
          if (ms.ix != ix) { ix = ms.ix; fx = (double)ix; }
          if (ms.iy != iy) { iy = ms.iy; fy = (double)iy; }
          if (ms.iz != iz) { iz = ms.iz; fz = (double)iz; }
          if (!in) state = LANE_ESCAPED;
        } else if ((uint32_t)ix >= ncx || (uint32_t)iy >= ncy || (uint32_t)iz >= ncz) {
          state = LANE_ESCAPED;
        }
        /* tau == 0 exactly: the walk ends inside (loop condition tau > 0, :391), on the wall */
        if (state == LANE_LIVE && !(tau > 0.)) state = LANE_ABSORBED;
        if (PRE && state == LANE_LIVE) prefetch_cell();
      }
    }
    if constexpr (AGG) {
      /* runs of neighbouring lanes with the same record (neighbours inside a group are neighbours
       * in key order): segmented sum towards the first lane of every run, log2(group) shuffle steps */
      const uint32_t key = do_acc ? akey : 0xffffffffu;
      const uint32_t prev = __shfl_up_sync(0xffffffffu, key, 1);
      const bool head = !do_acc || (lane & (AGG_GROUP - 1)) == 0 || key != prev;
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      if (heads != 0xffffffffu) {
        const unsigned above = heads & (0xfffffffeu << lane);
        const int last = above ? (__ffs(above) - 2) : 31;
        const bool heat = __any_sync(0xffffffffu, v[1] != 0.); /* a source at the threshold adds no heat */
#pragma unroll
        for (int d = 1; d < AGG_GROUP; d <<= 1) {
          const bool take = lane + d <= last;
#pragma unroll
          for (int k = 0; k < NV; ++k) {
            if (k == 1 && !heat) continue;
            const double t = __shfl_down_sync(0xffffffffu, v[k], d);
            if (take) v[k] += t;
          }
        }
      }
      const bool issue = do_acc && head;
      {
        if (issue) {
          if (v[0] != 0.) { atomicAdd(arec + (MODE == ACC_FULL ? acc_slot(ION_H_n) : 0), v[0]); ++n_red; }
          if (v[1] != 0.) { atomicAdd(arec + (MODE == ACC_FULL ? (int64_t)acc_slot(NUM_IONS + HEAT_H) : ats), v[1]); ++n_red; }
          if constexpr (MODE == ACC_FULL) {
            if (v[2] != 0.) { atomicAdd(arec + acc_slot(ION_O_n), v[2]); ++n_red; }
            if (v[3] != 0.) { atomicAdd(arec + acc_slot(ION_N_n), v[3]); ++n_red; }
          }
        }
      }
    }
  }
  ShootCounters cnt;
  cnt.n_steps = n_steps;
  cnt.n_red = n_red;
  cnt.tau_sum = tau_sum;
#pragma unroll
  for (int t = 0; t < NUM_PACKET_TYPES; ++t) {
    cnt.w_type[t] = (double)n_type[t] * w_discrete + (double)s_ntype_c[t][tid] * w_continuous;
    cnt.w_tot += cnt.w_type[t];
  }
  reduce_counters(P.acc, cnt);
}

} // namespace cmib
