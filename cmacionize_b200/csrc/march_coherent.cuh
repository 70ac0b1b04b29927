/*
 * march_coherent.cuh — the voxel walk of the H-only layout on the ORDERED march queue
 * (W.sort == 2), rewritten around the instruction budget of one cell crossing.
 *
 * Same contract as march_kernel (wavefront.cuh): CartesianDensityGrid::interact
 * (src/CartesianDensityGrid.cpp:375-452) + DensityGrid::update_integrals
 * (src/DensityGrid.hpp:150-197), one crossing per pass for every live lane, arithmetic of
 * march_step (march.cuh) operation for operation.  What changed, and why (ncu on
 * march_kernel<ACC_HONLY, AGG, PRE>, profiles/r01_coherent_march.md: 230 warp instructions per
 * pass, issue slots 72 % busy at 24 warps per SM, DRAM 24 %: the kernel is bound by instruction
 * issue, not by memory):
 *
 *  - cell walls come from shared-memory tables lo[i] = anchor + cellside * i and
 *    hi[i] = lo[i] + cellside (the reference's get_cell, :170-176, evaluated once per CTA with
 *    the same two roundings) instead of 3 x (DMUL + 2 DADD + select) per pass and three
 *    double-precision shadows of the cell indices;
 *  - per-packet constants that a crossing does not touch (id, meta, sampled tau, weight, frequency
 *    offset) live in shared memory; ds and tau_cell of an absorbed packet are parked there too,
 *    so that the loop fits in 64 registers = 4 CTAs (32 warps) per SM instead of 3;
 *  - the service test (finish + refill) runs once per LEAN_UNROLL crossings;
 *  - the heat term is a compile-time variant: a source at the ionisation threshold adds none
 *    (host side: monochromatic spectrum at nu_H and no re-emission);
 *  - hot-cell bookkeeping is a countdown that is zero after the first three crossings;
 *  - the accumulator address is formed by the run leaders only, after the in-warp sums.
 */
#pragma once
#include "wavefront.cuh"

namespace cmib {

#ifndef CMIB_LEAN_BLOCKS
#define CMIB_LEAN_BLOCKS 4
#endif
#ifndef CMIB_LEAN_UNROLL
#define CMIB_LEAN_UNROLL 2
#endif

/* explicit shared-memory accesses: a 32-bit shared address + immediate offset (one LDS / STS), so that
 * the per-thread base is ONE register the compiler cannot rematerialise inside the loop (it recomputed
 * SR_TID / SR_CgaCtaId based addresses at every use under the 64-register bound) */
template <int OFF> CMIB_D double lds_f64(uint32_t a) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF> CMIB_D void sts_f64(uint32_t a, double v) {
  asm volatile("st.shared.f64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "d"(v));
}
template <int OFF> CMIB_D unsigned long long lds_u64(uint32_t a) {
  unsigned long long v;
  asm volatile("ld.shared.u64 %0, [%1+%2];" : "=l"(v) : "r"(a), "n"(OFF));
  return v;
}
template <int OFF> CMIB_D void sts_u64(uint32_t a, unsigned long long v) {
  asm volatile("st.shared.u64 [%0+%1], %2;" ::"r"(a), "n"(OFF), "l"(v));
}
/* queue entries are read once, 8 bytes out of every 32-byte sector (the ordered queue is read through an index):
 * do not keep them in L1, and let L2 drop them first, so that the cell records and accumulators of the cone the
 * packets in flight walk through stay resident */
CMIB_D double ld_queue_f64(const double *p) {
  double v;
  asm volatile("{\n\t.reg .b64 pol;\n\tcreatepolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
               "ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], pol;\n\t}" : "=d"(v) : "l"(p));
  return v;
}
CMIB_D uint32_t ld_queue_u32(const uint32_t *p) {
  uint32_t v;
  asm volatile("{\n\t.reg .b64 pol;\n\tcreatepolicy.fractional.L2::evict_first.b64 pol, 1.0;\n\t"
               "ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], pol;\n\t}" : "=r"(v) : "l"(p));
  return v;
}
CMIB_D uint32_t opaque_u32(uint32_t v) {
  uint32_t r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

/* per-thread fields in shared memory: [field][MARCH_BLOCK] 8-byte slots */
enum LeanField : int {
  PT_ID = 0, PT_META, PT_TAU0, PT_W, PT_DNU, PT_DS, PT_TC, PT_TAUSUM, PT_SIGH,
  PT_NTYPE, /* NUM_PACKET_TYPES slots: ended packets of type t, discrete (low word) | continuous (high word) */
  PT_NFIELDS = PT_NTYPE + NUM_PACKET_TYPES
};
constexpr int PT_STRIDE = MARCH_BLOCK * 8;
constexpr int LEAN_PT_BYTES = PT_NFIELDS * PT_STRIDE;

/* dynamic shared memory: per-thread fields, then the wall tables (16 bytes per cell index and axis) */
inline size_t lean_smem_bytes(const GridGeom &g) {
  return (size_t)LEAN_PT_BYTES + (size_t)(g.ncell[0] + g.ncell[1] + g.ncell[2]) * 16;
}

/*
 * HEAT     the packets may carry nu != nu_H (heat term v1 = dJ_H * (nu - nu_H)).  false = the host established
 *          that every packet of the shoot is a primary of a discrete source at exactly nu_H (monochromatic
 *          spectrum at the threshold, no re-emission, no continuous source): then sigma_H and the weight are the
 *          same for all packets and come from the launch parameters (W.uni_sigH, W.uni_w) instead of two shared
 *          loads per crossing
 * PERIODIC some axis is periodic
 * PRE      request the record of the next cell at the end of a crossing
 * STEPS    shuffle steps of the in-warp sum: runs of up to 2^STEPS lanes are summed (<= 3)
 */
template <bool HEAT, bool PERIODIC, bool PRE, int STEPS>
__global__ void __launch_bounds__(MARCH_BLOCK, CMIB_LEAN_BLOCKS)
march_lean_kernel(const __grid_constant__ WavefrontParams W) {
  extern __shared__ double2 s_dyn[];
  __shared__ uint32_t s_hot_base;
  const ShootParams &P = W.sp;
  const GridGeom &g = P.geom;
  const uint64_t cap = W.capacity;
  const uint64_t qcount = W.ctl[CTL_QCOUNT];
  const int lane = threadIdx.x & 31;
  const bool can_reemit = (P.src.reemission_kind != REEMISSION_NONE);
  const uint32_t ncx = (uint32_t)g.ncell[0], ncy = (uint32_t)g.ncell[1], ncz = (uint32_t)g.ncell[2];
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(s_dyn);
  /* this thread's slot of field 0; field k is at + k * PT_STRIDE */
  const uint32_t pt = opaque_u32(smem0 + threadIdx.x * 8u);
  const uint32_t wall0 = opaque_u32(smem0 + (uint32_t)LEAN_PT_BYTES);
  {
    /* get_cell (CartesianDensityGrid.cpp:170-176): lo = anchor + cellside * i, hi = lo + cellside.  Entry k of an
     * axis holds (hi[k], lo[n - 1 - k]): a packet that moves up reads hi of its index i at entry i, one that
     * moves down reads lo at entry n - 1 - i, so that BOTH step to the next entry per crossing of the axis and
     * leave the grid at entry n */
    double2 *wt = s_dyn + LEAN_PT_BYTES / 16;
    for (uint32_t i = threadIdx.x; i < ncx + ncy + ncz; i += MARCH_BLOCK) {
      const int a = (i < ncx) ? 0 : (i < ncx + ncy ? 1 : 2);
      const uint32_t k = i - (a == 0 ? 0u : (a == 1 ? ncx : ncx + ncy));
      const uint32_t n = (a == 0) ? ncx : (a == 1 ? ncy : ncz);
      const double lo_up = xadd(g.anchor[a], xmul(g.cellside[a], (double)k));
      const double lo_dn = xadd(g.anchor[a], xmul(g.cellside[a], (double)(n - 1u - k)));
      wt[i] = make_double2(xadd(lo_up, g.cellside[a]), lo_dn);
    }
    unsigned long long *f = reinterpret_cast<unsigned long long *>(s_dyn);
#pragma unroll
    for (int k = 0; k < PT_NFIELDS; ++k) f[k * MARCH_BLOCK + threadIdx.x] = 0ull;
    if (threadIdx.x == 0) s_hot_base = (P.hot_replicas > 0) ? (blockIdx.x % P.hot_replicas) * HOT_MAX_SOURCES : 0u;
  }
  __syncthreads();

  /* packet state */
  double px = 0., py = 0., pz = 0., dx = 1., dy = 1., dz = 1., ivx = 1., ivy = 1., ivz = 1.;
  double tau = 0.;
  /* per axis: the shared address of the wall the packet moves towards (entry i, or entry n - 1 - i plus 8 bytes,
   * see above) and the step of the cell index per crossing of that axis (+-1, 0 never); the cell indices
   * themselves are only implicit */
  uint32_t ax = 0, ay = 0, az = 0;
  int32_t sx = 1, sy = 1, sz = 1;
  uint32_t cell = 0;
  /* table addresses of entry 0 and of entry n (= outside), per axis (uniform) */
  const uint32_t ux = wall0, uy = wall0 + W.lean_n16[0], uz = uy + W.lean_n16[1];
  const uint32_t ex = uy, ey = uz, ez = uz + W.lean_n16[2];
  const int32_t kx = W.lean_k[0], ky = W.lean_k[1]; /* ncy * ncz, ncz */
  uint32_t hot = 0; /* (source + 1) << 2 | crossings made; 0 = the neighbourhood rule no longer applies */
  uint32_t n_steps = 0, n_red = 0;
  int state = LANE_EMPTY;
  double pre_n = 0., pre_xH = 0.;
  uint64_t cur = 0, end = 0;
  bool exhausted = (qcount == 0);
  uint32_t n_pass = 0;
  const uint32_t cell_stride = (uint32_t)P.honly_cell_stride;

  while (true) {
    if (++n_pass > (1u << 26)) {
      if (lane == 0) atomicExch(&W.ctl[CTL_ERROR], 1ull);
      break;
    }
    const unsigned live_m = __ballot_sync(0xffffffffu, state == LANE_LIVE);
    const unsigned waiting = ~live_m;
    unsigned fill_m;
    {
      static_assert(AGG_GROUP == 8, "group mask arithmetic below is written for 8 lanes");
      unsigned m = waiting;
      m &= m >> 1;
      m &= m >> 2;
      m &= m >> 4;
      fill_m = (m & 0x01010101u) * 0xffu;
    }
    bool service = (fill_m != 0u);
    if (service && exhausted) {
      const unsigned pend_m = __ballot_sync(0xffffffffu, state == LANE_ABSORBED || state == LANE_ESCAPED);
      if (live_m == 0u && pend_m == 0u) break;
      service = (pend_m != 0u) && (__popc(pend_m) >= MARCH_REFILL_MIN || live_m == 0u);
    }
    if (service) {
      /* ---- finish ---- */
      double fpx = 0., fpy = 0., fpz = 0.;
      const unsigned long long meta_old = lds_u64<PT_META * PT_STRIDE>(pt);
      const uint32_t one_of_kind = meta_continuous(meta_old) ? 0x10000u : 1u; /* counters: two 16-bit... */
      int end_type = -1;
      if (state == LANE_ABSORBED && !(tau < 0.)) {
        /* tau == 0 exactly after a full crossing: the walk ends on the wall, in the cell the packet has
         * just entered (interact() returns the cell of the current index, :445-451); `cell` is that cell */
        fpx = px; fpy = py; fpz = pz;
        if (!can_reemit) end_type = PACKET_ABSORBED;
      } else if (state == LANE_ABSORBED) {
        /* tau < 0 after the crossing: shorten it (CartesianDensityGrid.cpp:413-417) */
        const double ds = lds_f64<PT_DS * PT_STRIDE>(pt), tau_cell = lds_f64<PT_TC * PT_STRIDE>(pt);
        const double nwx = xadd(px, xmul(ds, dx));
        const double nwy = xadd(py, xmul(ds, dy));
        const double nwz = xadd(pz, xmul(ds, dz));
        const double Scorr = xdiv(xmul(ds, tau), tau_cell);
        const double dss = xadd(ds, Scorr);
        fpx = xadd(px, xdiv(xmul(xsub(nwx, px), dss), ds));
        fpy = xadd(py, xdiv(xmul(xsub(nwy, py), dss), ds));
        fpz = xadd(pz, xdiv(xmul(xsub(nwz, pz), dss), ds));
        /* accumulate the shortened crossing; the cell has n > 0 (tau_cell > 0) */
        const double dJH = HEAT ? (dss * lds_f64<PT_W * PT_STRIDE>(pt)) * lds_f64<PT_SIGH * PT_STRIDE>(pt) : (dss * W.uni_w) * W.uni_sigH;
        if (dJH != 0.) {
          atomicAdd(acc_term<ACC_HONLY>(P, cell, 0), dJH);
          ++n_red;
          if (HEAT) {
            const double dh = dJH * lds_f64<PT_DNU * PT_STRIDE>(pt);
            if (dh != 0.) { atomicAdd(acc_term<ACC_HONLY>(P, cell, 1), dh); ++n_red; }
          }
        }
        if (!can_reemit) end_type = PACKET_ABSORBED; /* PhotonSource::reemit without a handler (:304-306) */
      } else if (state == LANE_ESCAPED) {
        end_type = meta_type(meta_old); /* keeps its last type (IonizationPhotonShootJob.hpp:143-144) */
      }
      if (end_type >= 0) {
        /* one 64-bit slot per type: discrete count in the low word, continuous in the high word */
        const uint32_t a = pt + (uint32_t)(PT_NTYPE + end_type) * PT_STRIDE;
        sts_u64<0>(a, lds_u64<0>(a) + (meta_continuous(meta_old) ? (1ull << 32) : 1ull));
      }
      (void)one_of_kind;
      if (state == LANE_ABSORBED || state == LANE_ESCAPED)
        sts_f64<PT_TAUSUM * PT_STRIDE>(pt, lds_f64<PT_TAUSUM * PT_STRIDE>(pt) +
                                               (lds_f64<PT_TAU0 * PT_STRIDE>(pt) - ((tau > 0.) ? tau : 0.)));
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
            __stcs(q + RQ_SIGH * cap, HEAT ? lds_f64<PT_SIGH * PT_STRIDE>(pt) : W.uni_sigH);
            __stcs(q + RQ_SIGHE * cap, 0.);
            __stcs(q + RQ_CELL * cap, __longlong_as_double((long long)cell));
            __stcs(q + RQ_ID * cap, __longlong_as_double((long long)lds_u64<PT_ID * PT_STRIDE>(pt)));
            __stcs(q + RQ_META * cap, __longlong_as_double((long long)(meta_old & 0xffffffffffull)));
          }
        }
      }
      if (state != LANE_LIVE) state = LANE_EMPTY;

      /* ---- refill: hand queue entries to the empty lanes of the waiting groups ---- */
      if (!exhausted) {
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
          const int rank = __popc(fill_m & ((1u << lane) - 1u));
          const uint64_t avail = end - cur;
          const int nfill = __popc(fill_m);
          if (state == LANE_EMPTY && ((fill_m >> lane) & 1u) && (uint64_t)rank < avail) {
            const uint64_t slot = (uint64_t)ld_queue_u32(W.order + cur + rank);
            const double *q = W.mq + slot;
            px = ld_queue_f64(q + MQ_PX * cap); py = ld_queue_f64(q + MQ_PY * cap); pz = ld_queue_f64(q + MQ_PZ * cap);
            dx = ld_queue_f64(q + MQ_DX * cap); dy = ld_queue_f64(q + MQ_DY * cap); dz = ld_queue_f64(q + MQ_DZ * cap);
            const double nu = ld_queue_f64(q + MQ_NU * cap);
            tau = ld_queue_f64(q + MQ_TAU * cap);
            sts_f64<PT_TAU0 * PT_STRIDE>(pt, tau);
            sts_u64<PT_ID * PT_STRIDE>(pt, (unsigned long long)__double_as_longlong(ld_queue_f64(q + MQ_ID * cap)));
            const unsigned long long meta = (unsigned long long)__double_as_longlong(ld_queue_f64(q + MQ_META * cap));
            sts_u64<PT_META * PT_STRIDE>(pt, meta);
            hot = 0;
            if (P.hot_replicas > 0) {
              const int isrc = meta_source(meta);
              if (isrc >= 0) hot = (uint32_t)(isrc + 1) << 2;
            }
            if (HEAT) {
              sts_f64<PT_SIGH * PT_STRIDE>(pt, ld_queue_f64(q + MQ_SIGMA * cap));
              sts_f64<PT_W * PT_STRIDE>(pt, meta_continuous(meta) ? P.src.continuous_weight : P.src.discrete_weight);
              sts_f64<PT_DNU * PT_STRIDE>(pt, nu - P.nu_H);
            }
            /* a zero direction component: the reference's wall distance is DBL_MAX (:289-309).  Here the
             * packet "moves towards" the upper wall with 1/d = +inf: (hi - p) * inf = +inf, which is never
             * the minimum and never equal to it — the same selection, without a test per crossing
             * (hi - p > 0: p lies in the cell of its truncated index and never moves along this axis) */
            const bool upx = (dx >= 0.), upy = (dy >= 0.), upz = (dz >= 0.);
            ivx = 1. / ((dx == 0.) ? 0. : dx);
            ivy = 1. / ((dy == 0.) ? 0. : dy);
            ivz = 1. / ((dz == 0.) ? 0. : dz);
            sx = upx ? 1 : -1;
            sy = upy ? 1 : -1;
            sz = upz ? 1 : -1;
            /* get_cell_indices (CartesianDensityGrid.cpp:152-161) */
            int32_t ix = trunc_index(xmul(xsub(px, g.anchor[0]), g.inv_cellside[0]));
            int32_t iy = trunc_index(xmul(xsub(py, g.anchor[1]), g.inv_cellside[1]));
            int32_t iz = trunc_index(xmul(xsub(pz, g.anchor[2]), g.inv_cellside[2]));
            state = LANE_LIVE;
            if (PERIODIC) {
              MarchState ms;
              ms.px = px; ms.py = py; ms.pz = pz; ms.ix = ix; ms.iy = iy; ms.iz = iz;
              const bool in = march_inside(g, ms);
              px = ms.px; py = ms.py; pz = ms.pz; ix = ms.ix; iy = ms.iy; iz = ms.iz;
              if (!in) state = LANE_ESCAPED;
            } else if ((uint32_t)ix >= ncx || (uint32_t)iy >= ncy || (uint32_t)iz >= ncz) {
              state = LANE_ESCAPED; /* emitted outside the box: interact() returns end() */
            }
            ax = ux + (upx ? 16u * (uint32_t)ix : 16u * (ncx - 1u - (uint32_t)ix) + 8u);
            ay = uy + (upy ? 16u * (uint32_t)iy : 16u * (ncy - 1u - (uint32_t)iy) + 8u);
            az = uz + (upz ? 16u * (uint32_t)iz : 16u * (ncz - 1u - (uint32_t)iz) + 8u);
            if (state == LANE_LIVE) {
              cell = ((uint32_t)ix * ncy + (uint32_t)iy) * ncz + (uint32_t)iz;
              if (PRE) {
                const double2 r0 = __ldg(P.cells_h + cell);
                pre_n = r0.x; pre_xH = r0.y;
              }
            }
          }
          cur += ((uint64_t)nfill < avail) ? (uint64_t)nfill : avail;
        }
      }
      continue;
    }

#pragma unroll
    for (int u = 0; u < CMIB_LEAN_UNROLL; ++u) {
      /* ---- one cell crossing for every live lane ---- */
      uint32_t akey = 0xffffffffu; /* accumulator record this lane adds to, as its index in doubles from W.acc_j (a
                                    * cell's J_H, or a hot replica record behind the cells); all ones = nothing to add */
      double v0 = 0., v1 = 0.;
      if (state == LANE_LIVE) {
        double n, xH;
        if (PRE) {
          n = pre_n; xH = pre_xH;
        } else {
          const double2 r0 = __ldg(P.cells_h + cell);
          n = r0.x; xH = r0.y;
        }
        /* get_wall_intersection (CartesianDensityGrid.cpp:280-318) on the tabulated walls */
        const double wx = xmul(xsub(lds_f64<0>(ax), px), ivx);
        const double wy = xmul(xsub(lds_f64<0>(ay), py), ivy);
        const double wz = xmul(xsub(lds_f64<0>(az), pz), ivz);
        const double sigH = HEAT ? lds_f64<PT_SIGH * PT_STRIDE>(pt) : W.uni_sigH;
        const double myz = (wz < wy) ? wz : wy;
        const double ds = (myz < wx) ? myz : wx;
        /* ds * n * (sigma_H * x_H) (DensityGrid.hpp:129-133; the helium term of the H-only layout is
         * sigma * 0 + 0 = +0, and t + 0 == t) */
        const double tau_cell = xmul(xmul(ds, n), xmul(sigH, xH));
        tau = xsub(tau, tau_cell);
        ++n_steps;
        if (tau < 0.) {
          state = LANE_ABSORBED; /* position and tau stay as they are for the deferred finish */
          sts_f64<PT_DS * PT_STRIDE>(pt, ds);
          sts_f64<PT_TC * PT_STRIDE>(pt, tau_cell);
        } else {
          if (n > 0.) {
            /* update_integrals (DensityGrid.hpp:150-197) */
            akey = cell * cell_stride;
            if (hot != 0u) {
              /* first crossings of a primary, inside the 3x3x3 cells around its source: a replica */
              const uint32_t hc = P.src_cell[(hot >> 2) - 1u];
              const int jx = (int)((ax - ux) >> 4), jy = (int)((ay - uy) >> 4), jz = (int)((az - uz) >> 4);
              const int ddx = (sx > 0 ? jx : (int)ncx - 1 - jx) - (int)(hc & 1023u),
                        ddy = (sy > 0 ? jy : (int)ncy - 1 - jy) - (int)((hc >> 10) & 1023u),
                        ddz = (sz > 0 ? jz : (int)ncz - 1 - jz) - (int)((hc >> 20) & 1023u);
              if ((unsigned)(ddx + 1) < 3u && (unsigned)(ddy + 1) < 3u && (unsigned)(ddz + 1) < 3u)
                akey = W.hot_index0 + (uint32_t)HOT_STRIDE * ((s_hot_base + ((hot >> 2) - 1u)) * HOT_CELLS +
                                                              (uint32_t)((ddx + 1) * 9 + (ddy + 1) * 3 + (ddz + 1)));
              ++hot;
              if ((hot & 3u) == (uint32_t)HOT_CROSSINGS) hot = 0u;
            }
            v0 = (ds * (HEAT ? lds_f64<PT_W * PT_STRIDE>(pt) : W.uni_w)) * sigH;
            if (HEAT) v1 = v0 * lds_f64<PT_DNU * PT_STRIDE>(pt);
          }
          /* move to the wall, step the indices of every axis whose wall was hit */
          px = xadd(px, xmul(ds, dx));
          py = xadd(py, xmul(ds, dy));
          pz = xadd(pz, xmul(ds, dz));
          /* if (w == ds) { a += 16; cell += k * s; } as one compare and two predicated integer operations */
          asm("{\n\t.reg .pred p;\n\tsetp.eq.f64 p, %2, %3;\n\t@p add.u32 %0, %0, 16;\n\t@p mad.lo.s32 %1, %4, %5, %1;\n\t}"
              : "+r"(ax), "+r"(cell) : "d"(wx), "d"(ds), "r"(sx), "r"(kx));
          asm("{\n\t.reg .pred p;\n\tsetp.eq.f64 p, %2, %3;\n\t@p add.u32 %0, %0, 16;\n\t@p mad.lo.s32 %1, %4, %5, %1;\n\t}"
              : "+r"(ay), "+r"(cell) : "d"(wy), "d"(ds), "r"(sy), "r"(ky));
          asm("{\n\t.reg .pred p;\n\tsetp.eq.f64 p, %2, %3;\n\t@p add.u32 %0, %0, 16;\n\t@p add.s32 %1, %1, %4;\n\t}"
              : "+r"(az), "+r"(cell) : "d"(wz), "d"(ds), "r"(sz));
          if (PERIODIC) {
            /* is_inside (:187-227): an index that left a periodic axis re-enters on the other side, the position
             * moves by the box side; rare (once per box crossing), so the indices are recovered from the addresses */
            if (ax >= ex || ay >= ey || az >= ez) {
              MarchState ms;
              ms.px = px; ms.py = py; ms.pz = pz;
              const int32_t jx = (int32_t)((ax - ux) >> 4), jy = (int32_t)((ay - uy) >> 4), jz = (int32_t)((az - uz) >> 4);
              ms.ix = (sx > 0) ? jx : (int32_t)ncx - 1 - jx;
              ms.iy = (sy > 0) ? jy : (int32_t)ncy - 1 - jy;
              ms.iz = (sz > 0) ? jz : (int32_t)ncz - 1 - jz;
              const bool in = march_inside(g, ms);
              px = ms.px; py = ms.py; pz = ms.pz;
              ax = ux + ((sx > 0) ? 16u * (uint32_t)ms.ix : 16u * (ncx - 1u - (uint32_t)ms.ix) + 8u);
              ay = uy + ((sy > 0) ? 16u * (uint32_t)ms.iy : 16u * (ncy - 1u - (uint32_t)ms.iy) + 8u);
              az = uz + ((sz > 0) ? 16u * (uint32_t)ms.iz : 16u * (ncz - 1u - (uint32_t)ms.iz) + 8u);
              cell = ((uint32_t)ms.ix * ncy + (uint32_t)ms.iy) * ncz + (uint32_t)ms.iz;
              if (!in) state = LANE_ESCAPED;
            }
          } else if (ax >= ex || ay >= ey || az >= ez) {
            state = LANE_ESCAPED; /* entry n of an axis: the index left the grid */
          }
          if (state == LANE_LIVE) {
            /* tau == 0 exactly: the walk ends inside (loop condition tau > 0, :391), on the wall */
            if (!(tau > 0.)) state = LANE_ABSORBED;
            else if (PRE) {
              const double2 r0 = __ldg(P.cells_h + cell);
              pre_n = r0.x; pre_xH = r0.y;
            }
          }
        }
      }
      /* runs of neighbouring lanes with the same record (neighbours inside a group are neighbours
       * in key order): segmented sum towards the first lane of every run */
      bool head;
      {
        /* head = first lane of its group of 2^STEPS lanes, or a record different from the lane below: the shuffle's
         * own predicate says whether the lane below belongs to the same group (c = segment mask | clamp) */
        uint32_t prev;
        int in_group;
        asm volatile("{\n\t.reg .pred p;\n\tshfl.sync.up.b32 %0|p, %2, 1, %3, 0xffffffff;\n\tselp.s32 %1, 1, 0, p;\n\t}"
                     : "=r"(prev), "=r"(in_group) : "r"(akey), "n"((32 - (1 << STEPS)) << 8));
        head = !in_group || akey != prev;
      }
      const unsigned heads = __ballot_sync(0xffffffffu, head);
      if (heads != 0xffffffffu) {
        /* last lane of my run = lane before the next head above me (lane 32 counts as a head); a shuffle from
         * beyond it is refused by the clamp operand and its predicate gates the add */
        const unsigned hs = ((heads >> 1) | 0x80000000u) >> lane;
        const uint32_t run_last = (uint32_t)lane + (uint32_t)(__ffs(hs) - 1);
#pragma unroll
        for (int d = 1; d < (1 << STEPS); d <<= 1) {
          asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, tl, th;\n\t.reg .f64 t;\n\t"
                       "mov.b64 {lo, hi}, %0;\n\t"
                       "shfl.sync.down.b32 tl|p, lo, %1, %2, 0xffffffff;\n\t"
                       "shfl.sync.down.b32 th, hi, %1, %2, 0xffffffff;\n\t"
                       "mov.b64 t, {tl, th};\n\t"
                       "@p add.rn.f64 %0, %0, t;\n\t}"
                       : "+d"(v0) : "r"(d), "r"(run_last));
          if (HEAT) {
            asm volatile("{\n\t.reg .pred p;\n\t.reg .b32 lo, hi, tl, th;\n\t.reg .f64 t;\n\t"
                         "mov.b64 {lo, hi}, %0;\n\t"
                         "shfl.sync.down.b32 tl|p, lo, %1, %2, 0xffffffff;\n\t"
                         "shfl.sync.down.b32 th, hi, %1, %2, 0xffffffff;\n\t"
                         "mov.b64 t, {tl, th};\n\t"
                         "@p add.rn.f64 %0, %0, t;\n\t}"
                         : "+d"(v1) : "r"(d), "r"(run_last));
          }
        }
      }
      if (head && akey != 0xffffffffu) {
        double *a = W.acc_j + akey;
        if (v0 != 0.) { atomicAdd(a, v0); ++n_red; }
        if (HEAT && v1 != 0.) { atomicAdd(a + ((akey >= W.hot_index0) ? (int64_t)1 : P.honly_term_stride), v1); ++n_red; }
      }
    }
  }
  ShootCounters cnt;
  cnt.n_steps = n_steps;
  cnt.n_red = n_red;
  cnt.tau_sum = lds_f64<PT_TAUSUM * PT_STRIDE>(pt);
  const double w_discrete = P.src.discrete_weight, w_continuous = P.src.continuous_weight;
#pragma unroll
  for (int t = 0; t < NUM_PACKET_TYPES; ++t) {
    const unsigned long long c = lds_u64<0>(pt + (uint32_t)(PT_NTYPE + t) * PT_STRIDE);
    cnt.w_type[t] = (double)(uint32_t)c * w_discrete + (double)(uint32_t)(c >> 32) * w_continuous;
    cnt.w_tot += cnt.w_type[t];
  }
  reduce_counters(P.acc, cnt);
}

} // namespace cmib
