/*
 * march.cuh — voxel walk of one photon packet through the Cartesian grid.
 *
 * Behavioural contract = CartesianDensityGrid::interact
 * (/root/reference/src/CartesianDensityGrid.cpp:375-452) with
 *   get_cell_indices      :152-161   (truncating start index)
 *   is_inside             :187-227   (periodic wrap mutates index AND position)
 *   get_cell              :170-176   (lo = anchor + cellside*i ; hi = lo + cellside)
 *   get_wall_intersection :280-318   (equality-based face selection; ties step
 *                                     several axes at once)
 *   get_optical_depth     DensityGrid.hpp:117-140
 *   update_integrals      DensityGrid.hpp:150-197
 * and the operation order spelled out in SURVEY.md Appendix A.  All arithmetic
 * that feeds a comparison or an index goes through xmul/xadd/xsub/xdiv (no FMA
 * contraction) so that the visited-cell sequence is bit-identical.
 *
 * The walk is written as a resumable state machine (init / step) so that the
 * shoot kernel can interleave refill of finished lanes with marching.
 */
#pragma once
#include "cmib_common.cuh"

namespace cmib {

struct MarchState {
  double px, py, pz;    /* position (m) */
  double dx, dy, dz;    /* direction */
  double ix_, iy_, iz_; /* inverse direction (1/d, may be inf) */
  double tau;           /* optical depth still to travel */
  int32_t ix, iy, iz;   /* current cell indices (may be outside) */
  int64_t last_cell;    /* long index of the last cell a step was taken in, -1 if none */
};

CMIB_HD int64_t long_index(const GridGeom &g, int32_t ix, int32_t iy, int32_t iz) {
  /* CartesianDensityGrid.hpp:137-144: ix*ny*nz + iy*nz + iz */
  return ((int64_t)ix * g.ncell[1] + iy) * (int64_t)g.ncell[2] + iz;
}

CMIB_HD int32_t trunc_index(double v) {
  /* C++ double -> int_fast32_t conversion truncates toward zero (Appendix A.2) */
  /* the reference holds the index in a 64-bit int_fast32_t; clamp so that an
   * absurdly distant position cannot alias into the box after narrowing */
#if defined(__CUDA_ARCH__)
  long long i = __double2ll_rz(v);
#else
  long long i = (v != v) ? 0 : (v > 4.e18 ? (1ll << 62) : (v < -4.e18 ? -(1ll << 62) : (long long)v));
#endif
  if (i > (1ll << 30)) i = (1ll << 30);
  if (i < -(1ll << 30)) i = -(1ll << 30);
  return (int32_t)i;
}

CMIB_HD void march_locate(const GridGeom &g, MarchState &s) {
  s.ix = trunc_index(xmul(xsub(s.px, g.anchor[0]), g.inv_cellside[0]));
  s.iy = trunc_index(xmul(xsub(s.py, g.anchor[1]), g.inv_cellside[1]));
  s.iz = trunc_index(xmul(xsub(s.pz, g.anchor[2]), g.inv_cellside[2]));
  s.last_cell = -1;
}

/* is_inside: returns validity, applies the periodic wrap (index and position) */
CMIB_HD bool march_inside(const GridGeom &g, MarchState &s) {
  bool inside = true;
  if (!g.periodic[0]) {
    inside &= (s.ix >= 0 && s.ix < g.ncell[0]);
  } else {
    if (s.ix < 0) { s.ix = g.ncell[0] - 1; s.px = xadd(s.px, g.sides[0]); }
    if (s.ix >= g.ncell[0]) { s.ix = 0; s.px = xsub(s.px, g.sides[0]); }
  }
  if (!g.periodic[1]) {
    inside &= (s.iy >= 0 && s.iy < g.ncell[1]);
  } else {
    if (s.iy < 0) { s.iy = g.ncell[1] - 1; s.py = xadd(s.py, g.sides[1]); }
    if (s.iy >= g.ncell[1]) { s.iy = 0; s.py = xsub(s.py, g.sides[1]); }
  }
  if (!g.periodic[2]) {
    inside &= (s.iz >= 0 && s.iz < g.ncell[2]);
  } else {
    if (s.iz < 0) { s.iz = g.ncell[2] - 1; s.pz = xadd(s.pz, g.sides[2]); }
    if (s.iz >= g.ncell[2]) { s.iz = 0; s.pz = xsub(s.pz, g.sides[2]); }
  }
  return inside;
}

CMIB_HD double wall_distance(double lo, double cs, double p, double d, double id) {
  /* get_wall_intersection :289-309; hi = lo + cellside (Box::get_top_anchor) */
  if (d > 0.) return xmul(xsub(xadd(lo, cs), p), id);
  if (d < 0.) return xmul(xsub(lo, p), id);
  return DBL_MAX;
}

/*
 * One cell crossing.  Preconditions: march_inside(g,s) returned true and
 * s.tau > 0.  `n`, `xH`, `xHe` are the cell's values, sigH / sigHe the packet's
 * sigma_H and A_He*sigma_He.  Returns the path length actually travelled in the
 * cell (possibly shortened when the optical depth ran out); s is advanced.
 */
CMIB_HD double march_step(const GridGeom &g, MarchState &s, double n, double xH, double xHe,
                          double sigH, double sigHe) {
  const double lox = xadd(g.anchor[0], xmul(g.cellside[0], (double)s.ix));
  const double loy = xadd(g.anchor[1], xmul(g.cellside[1], (double)s.iy));
  const double loz = xadd(g.anchor[2], xmul(g.cellside[2], (double)s.iz));
  const double wx = wall_distance(lox, g.cellside[0], s.px, s.dx, s.ix_);
  const double wy = wall_distance(loy, g.cellside[1], s.py, s.dy, s.iy_);
  const double wz = wall_distance(loz, g.cellside[2], s.pz, s.dz, s.iz_);
  /* std::min(dx, std::min(dy, dz)) */
  const double myz = (wz < wy) ? wz : wy;
  double ds = (myz < wx) ? myz : wx;
  const int32_t nx = (wx == ds) ? ((s.dx > 0.) ? 1 : -1) : 0;
  const int32_t ny = (wy == ds) ? ((s.dy > 0.) ? 1 : -1) : 0;
  const int32_t nz = (wz == ds) ? ((s.dz > 0.) ? 1 : -1) : 0;
  const double nwx = xadd(s.px, xmul(ds, s.dx));
  const double nwy = xadd(s.py, xmul(ds, s.dy));
  const double nwz = xadd(s.pz, xmul(ds, s.dz));
  /* ds * n * (sigma_H*x_H + sigma_Hecorr*x_He), left to right */
  const double tau_cell =
      xmul(xmul(ds, n), xadd(xmul(sigH, xH), xmul(sigHe, xHe)));
  s.tau = xsub(s.tau, tau_cell);
  if (s.tau < 0.) {
    const double Scorr = xdiv(xmul(ds, s.tau), tau_cell);
    const double dss = xadd(ds, Scorr);
    s.px = xadd(s.px, xdiv(xmul(xsub(nwx, s.px), dss), ds));
    s.py = xadd(s.py, xdiv(xmul(xsub(nwy, s.py), dss), ds));
    s.pz = xadd(s.pz, xdiv(xmul(xsub(nwz, s.pz), dss), ds));
    ds = dss;
  } else {
    s.px = nwx; s.py = nwy; s.pz = nwz;
    s.ix += nx; s.iy += ny; s.iz += nz;
  }
  return ds;
}

/*
 * CartesianDensityGrid::integrate_optical_depth (CartesianDensityGrid.cpp:328-363): optical depth
 * from the packet's position to the edge of the box, summed crossing by crossing in the
 * reference's order (get_optical_depth per cell, DensityGrid.hpp:117-140).  `cell_at(long index)`
 * returns the cell record.  A direction along a periodic axis never leaves the box — the reference
 * loops forever there — so the walk stops after max_crossings and returns what it has.
 */
template <class CellAt>
CMIB_HD double integrate_optical_depth(const GridGeom &g, MarchState s, double sigH, double sigHe, const CellAt &cell_at,
                                       int64_t max_crossings) {
  double optical_depth = 0.;
  s.ix_ = 1. / s.dx;
  s.iy_ = 1. / s.dy;
  s.iz_ = 1. / s.dz;
  s.tau = DBL_MAX; /* never exhausted: march_step always goes to the wall */
  march_locate(g, s);
  for (int64_t k = 0; k < max_crossings && march_inside(g, s); ++k) {
    const CellOpacity c = cell_at(long_index(g, s.ix, s.iy, s.iz));
    const double ds = march_step(g, s, c.n, c.xH, c.xHe, sigH, sigHe);
    s.tau = DBL_MAX;
    optical_depth = xadd(optical_depth, xmul(xmul(ds, c.n), xadd(xmul(sigH, c.xH), xmul(sigHe, c.xHe))));
  }
  return optical_depth;
}

} // namespace cmib
