/*
 * SPHArrayInterface.hpp — coupling of the Cartesian GPU grid to the particle arrays of an SPH
 * code: how SWIFT / PHANTOM-style hydro codes call CMacIonize (c/cmi_c_library.h).
 *
 * Behavioural contract = SPHArrayInterface (/root/reference/src/SPHArrayInterface.cpp, .hpp) for
 * the two mappings that make sense on a Cartesian grid:
 *   "M_over_V"  reset :147-189, operator() :941-942 (density = m[0] / cell volume: equal-mass
 *               particles, one per cell volume), inverse mapping .hpp:149-155 (the particle closest
 *               to a cell midpoint takes that cell's neutral fraction; Octree::get_closest_ngb)
 *   "centroid"  operator() :943-959 (SPH density at the cell midpoint: sum of m W(r/h, h) over the
 *               particles whose smoothing sphere contains the midpoint, Octree::get_ngbs :129-162;
 *               cubic spline CubicSplineKernel.hpp:44-59), inverse mapping .hpp:156-199 (every
 *               particle starts neutral and loses splineval / cell_mass * (1 - x_H(cell)))
 *   "Petkova"   needs the analytic kernel-volume integrals over Voronoi faces (:286-923): rejected.
 * Floor (:1006-1008): a cell no particle reaches gets m[0] / V * 1e-6.  mass -> number density with
 * the reference's hydrogen mass 1.6737236e-27 kg (:1010), T = 8000 K, x_H = x_He = 1e-6.
 *
 * The reference finds neighbours with an octree; here particles are scattered over the cells their
 * smoothing sphere touches, in particle order.  Sums therefore differ from the reference's in the
 * order of the additions only (tests: <= 1e-13 relative).  Quirk kept: after reset() the
 * non-periodic interface holds the particles' bounding box, so its kernel distances go through
 * Box::periodic_distance with those sides (.cpp:949-952) even though the neighbour search does not.
 */
#pragma once
#include <cfloat>
#include <cmath>
#include <string>
#include <vector>

#include "IonizationSimulation.hpp"

namespace cmi {

enum SPHArrayMappingType { SPHARRAY_MAPPING_M_OVER_V = 0, SPHARRAY_MAPPING_CENTROID, SPHARRAY_MAPPING_PETKOVA };

class SPHArrayInterface : public DensityFunction {
public:
  SPHArrayInterface(double unit_length_in_SI, double unit_mass_in_SI, const std::string &mapping_type)
      : unit_length_(unit_length_in_SI), unit_mass_(unit_mass_in_SI), is_periodic_(false),
        mapping_type_(get_mapping_type(mapping_type)) {
    box_anchor_ = {0., 0., 0.};
    box_sides_ = {0., 0., 0.};
  }
  template <typename T>
  SPHArrayInterface(double unit_length_in_SI, double unit_mass_in_SI, const T *box_anchor, const T *box_sides,
                    const std::string &mapping_type)
      : unit_length_(unit_length_in_SI), unit_mass_(unit_mass_in_SI), is_periodic_(true),
        mapping_type_(get_mapping_type(mapping_type)) {
    for (int d = 0; d < 3; ++d) {
      box_anchor_[d] = box_anchor[d] * unit_length_;
      box_sides_[d] = box_sides[d] * unit_length_;
    }
  }

  static SPHArrayMappingType get_mapping_type(const std::string &name) {
    if (name == "M_over_V") return SPHARRAY_MAPPING_M_OVER_V;
    if (name == "centroid") return SPHARRAY_MAPPING_CENTROID;
    if (name == "Petkova")
      cmi_error("SPHArrayMappingType \"Petkova\" is not provided by the B200 backend (Cartesian grids: M_over_V, centroid)!");
    cmi_error("Unknown SPHArrayMappingType: \"%s\"!", name.c_str());
  }

  /* reset (:147-270): copy the arrays in SI units; non-periodic: box = bounding box of the particles,
   * anchor - 0.5 %, sides + 1 % */
  template <typename TX, typename TH>
  void reset(const TX *x, const TX *y, const TX *z, const TH *h, const TH *m, size_t npart) {
    positions_.resize(npart);
    smoothing_lengths_.assign(npart, 0.);
    masses_.assign(npart, 0.);
    neutral_fractions_.assign(npart, 0.);
    for (size_t i = 0; i < npart; ++i) {
      positions_[i] = {x[i] * unit_length_, y[i] * unit_length_, z[i] * unit_length_};
      smoothing_lengths_[i] = h[i] * unit_length_;
      masses_[i] = m[i] * unit_mass_;
    }
    if (!is_periodic_) {
      Vec3 minpos = {DBL_MAX, DBL_MAX, DBL_MAX}, maxpos = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
      for (size_t i = 0; i < npart; ++i)
        for (int d = 0; d < 3; ++d) {
          minpos[d] = std::min(minpos[d], positions_[i][d]);
          maxpos[d] = std::max(maxpos[d], positions_[i][d]);
        }
      for (int d = 0; d < 3; ++d) {
        maxpos[d] -= minpos[d];
        box_anchor_[d] = minpos[d] - 0.005 * maxpos[d];
        box_sides_[d] = 1.01 * maxpos[d];
      }
    }
    cell_mass_density_.clear();
  }

  /* DensityFunction interface: the whole grid is filled at once (scatter over particles) */
  DensityValues operator()(const Vec3 &) override { cmi_error("SPHArrayInterface fills the grid through set_densities()!"); }
  bool set_densities(CartesianCells &grid) override {
    map_densities(grid);
    return true;
  }

  /* the DensityFunction stage (operator() for every cell, :932-1018), written into the grid's host mirror */
  void map_densities(CartesianCells &grid) {
    const size_t ncell = grid.get_number_of_cells();
    const double volume = grid.get_cell_volume();
    if (masses_.empty()) cmi_error("SPHArrayInterface: no particles!");
    cell_mass_density_.assign(ncell, 0.);
    if (mapping_type_ == SPHARRAY_MAPPING_M_OVER_V) {
      for (size_t c = 0; c < ncell; ++c) cell_mass_density_[c] = masses_[0] / volume;
    } else {
      for_each_particle_cell(grid, [&](size_t i, size_t c, double splineval) {
        (void)i;
        cell_mass_density_[c] += splineval;
      });
    }
    for (size_t c = 0; c < ncell; ++c) {
      double density = cell_mass_density_[c];
      if (density <= 0.0) density = masses_[0] / volume * 1e-6;
      grid.number_density[c] = density / 1.6737236e-27;
      grid.temperature[c] = 8000.;
      for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) grid.ionic_fraction[(size_t)ion * ncell + c] = 0.;
      grid.ionic_fraction[c] = 1.e-6;
      grid.ionic_fraction[ncell + c] = 1.e-6;
    }
  }

  /* DensityGridWriter::write = the inverse mapping (.cpp:1048-1075, .hpp:147-199); grid.ionic_fraction
   * must hold the result of the run (download() first) */
  void write(const CartesianCells &grid) {
    const size_t ncell = grid.get_number_of_cells();
    neutral_fractions_.assign(positions_.size(), 1.0);
    if (mapping_type_ == SPHARRAY_MAPPING_M_OVER_V) {
      build_bins();
      /* cells in index order; several cells can share a closest particle: the last one wins, as in
       * a single-threaded reference run */
      for (size_t c = 0; c < ncell; ++c) neutral_fractions_[closest_particle(grid.get_cell_midpoint(c))] = grid.ionic_fraction[c];
      return;
    }
    if (cell_mass_density_.size() != ncell) cmi_error("SPHArrayInterface::write before map_densities!");
    for_each_particle_cell(grid, [&](size_t i, size_t c, double splineval) {
      neutral_fractions_[i] -= splineval / cell_mass_density_[c] * (1. - grid.ionic_fraction[c]);
    });
  }

  template <typename T> void fill_array(T *nH) const {
    for (size_t i = 0; i < neutral_fractions_.size(); ++i) nH[i] = (T)neutral_fractions_[i];
  }
  size_t get_number_of_particles() const { return positions_.size(); }

private:
  /* Box::periodic_distance (src/Box.hpp:114-127) with this interface's box */
  Vec3 box_distance(const Vec3 &a, const Vec3 &b) const {
    Vec3 c = {a[0] - b[0], a[1] - b[1], a[2] - b[2]};
    for (int d = 0; d < 3; ++d) {
      if (2 * c[d] < -box_sides_[d]) c[d] += box_sides_[d];
      if (2 * c[d] >= box_sides_[d]) c[d] -= box_sides_[d];
    }
    return c;
  }
  static double norm(const Vec3 &v) { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
  /* CubicSplineKernel::kernel_evaluate (src/CubicSplineKernel.hpp:44-59) */
  static double kernel_evaluate(double u, double h) {
    const double KC1 = 2.546479089470, KC2 = 15.278874536822, KC5 = 5.092958178941;
    if (u < 1.) {
      if (u < 0.5) return (KC1 + KC2 * (u - 1.) * u * u) / (h * h * h);
      return KC5 * (1. - u) * (1. - u) * (1. - u) / (h * h * h);
    }
    return 0.;
  }

  /* f(particle, cell, m W) for every (particle, cell) pair the reference's get_ngbs(cell midpoint)
   * would return: distance (periodic when the interface is periodic) <= h of the particle */
  template <class F> void for_each_particle_cell(const CartesianCells &grid, F f) const {
    const std::array<int32_t, 3> &nc = grid.get_number_of_cells_3d();
    const Vec3 m0 = grid.get_cell_midpoint(0);
    const size_t last = grid.get_number_of_cells() - 1;
    const Vec3 m1 = grid.get_cell_midpoint(last);
    Vec3 cs;
    for (int d = 0; d < 3; ++d) cs[d] = (nc[d] > 1) ? (m1[d] - m0[d]) / (nc[d] - 1) : 1.;
    for (size_t i = 0; i < positions_.size(); ++i) {
      const Vec3 &p = positions_[i];
      const double h = smoothing_lengths_[i], m = masses_[i];
      int lo[3], hi[3];
      for (int d = 0; d < 3; ++d) {
        lo[d] = (int)std::floor((p[d] - h - m0[d]) / cs[d]) - 1;
        hi[d] = (int)std::ceil((p[d] + h - m0[d]) / cs[d]) + 1;
        if (!is_periodic_) {
          lo[d] = std::max(lo[d], 0);
          hi[d] = std::min(hi[d], nc[d] - 1);
        } else if (hi[d] - lo[d] >= nc[d]) {
          lo[d] = 0;
          hi[d] = nc[d] - 1;
        }
      }
      for (int ix = lo[0]; ix <= hi[0]; ++ix)
        for (int iy = lo[1]; iy <= hi[1]; ++iy)
          for (int iz = lo[2]; iz <= hi[2]; ++iz) {
            const int wx = ((ix % nc[0]) + nc[0]) % nc[0], wy = ((iy % nc[1]) + nc[1]) % nc[1],
                      wz = ((iz % nc[2]) + nc[2]) % nc[2];
            const size_t c = ((size_t)wx * nc[1] + wy) * nc[2] + wz;
            const Vec3 mid = grid.get_cell_midpoint(c);
            /* neighbour test of the octree: periodic distance only for a periodic interface (Octree.hpp:136-144) */
            const double rsel = is_periodic_ ? norm(box_distance(p, mid)) : norm({p[0] - mid[0], p[1] - mid[1], p[2] - mid[2]});
            if (!(rsel <= h)) continue;
            /* kernel distance (.cpp:946-952): through the box whenever it has a size */
            const double r = (!box_sides_[0]) ? norm({mid[0] - p[0], mid[1] - p[1], mid[2] - p[2]}) : norm(box_distance(mid, p));
            f(i, c, m * kernel_evaluate(r / h, h));
          }
    }
  }

  /* Octree::get_closest_ngb (src/Octree.hpp:279-315).  Particles are binned on a regular mesh over the
   * interface's box; bins are visited in shells of growing Chebyshev distance around the query until
   * no unvisited shell can hold a closer particle.  Exact ties are resolved towards the larger
   * particle index (the reference: the later one in tree order). */
  void build_bins() {
    const size_t n = positions_.size();
    nbin_ = std::max(1, (int)std::cbrt((double)n / 4.));
    bin_start_.assign((size_t)nbin_ * nbin_ * nbin_ + 1, 0);
    bin_items_.resize(n);
    std::vector<uint32_t> which(n);
    for (size_t i = 0; i < n; ++i) {
      which[i] = bin_of(positions_[i]);
      ++bin_start_[which[i] + 1];
    }
    for (size_t b = 1; b < bin_start_.size(); ++b) bin_start_[b] += bin_start_[b - 1];
    std::vector<uint32_t> fill(bin_start_.begin(), bin_start_.end() - 1);
    for (size_t i = 0; i < n; ++i) bin_items_[fill[which[i]]++] = (uint32_t)i;
  }
  void bin_index(const Vec3 &p, int *b) const {
    for (int d = 0; d < 3; ++d) {
      const double f = (box_sides_[d] > 0.) ? (p[d] - box_anchor_[d]) / box_sides_[d] * nbin_ : 0.;
      b[d] = std::min(nbin_ - 1, std::max(0, (int)std::floor(f)));
    }
  }
  uint32_t bin_of(const Vec3 &p) const {
    int b[3];
    bin_index(p, b);
    return (uint32_t)((b[0] * nbin_ + b[1]) * nbin_ + b[2]);
  }
  size_t closest_particle(const Vec3 &mid) const {
    int b0[3];
    bin_index(mid, b0);
    const double binsize = std::min(box_sides_[0], std::min(box_sides_[1], box_sides_[2])) / nbin_;
    double rmin = DBL_MAX;
    size_t imin = 0;
    for (int shell = 0; shell < nbin_ + 1; ++shell) {
      if (rmin < DBL_MAX && (shell - 1) * binsize > rmin) break;
      for (int dx = -shell; dx <= shell; ++dx)
        for (int dy = -shell; dy <= shell; ++dy)
          for (int dz = -shell; dz <= shell; ++dz) {
            if (std::max(std::abs(dx), std::max(std::abs(dy), std::abs(dz))) != shell) continue;
            int b[3] = {b0[0] + dx, b0[1] + dy, b0[2] + dz};
            bool skip = false;
            for (int d = 0; d < 3; ++d) {
              if (is_periodic_) {
                b[d] = ((b[d] % nbin_) + nbin_) % nbin_; /* wide shells revisit bins through the wrap: harmless */
              } else if (b[d] < 0 || b[d] >= nbin_) {
                skip = true;
              }
            }
            if (skip) continue;
            const size_t bin = ((size_t)b[0] * nbin_ + b[1]) * nbin_ + b[2];
            for (uint32_t k = bin_start_[bin]; k < bin_start_[bin + 1]; ++k) {
              const size_t i = bin_items_[k];
              const Vec3 &p = positions_[i];
              const double r = is_periodic_ ? norm(box_distance(p, mid)) : norm({p[0] - mid[0], p[1] - mid[1], p[2] - mid[2]});
              if (r < rmin || (r == rmin && i > imin)) {
                rmin = r;
                imin = i;
              }
            }
          }
    }
    return imin;
  }

  double unit_length_, unit_mass_;
  bool is_periodic_;
  SPHArrayMappingType mapping_type_;
  Vec3 box_anchor_, box_sides_;
  std::vector<Vec3> positions_;
  std::vector<double> smoothing_lengths_, masses_, neutral_fractions_;
  std::vector<double> cell_mass_density_; /* sum of m W per cell (centroid) */
  int nbin_ = 1;                          /* closest-particle search mesh */
  std::vector<uint32_t> bin_start_, bin_items_;
};

} // namespace cmi
