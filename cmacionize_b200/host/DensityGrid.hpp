/*
 * DensityGrid.hpp — CartesianCells (geometry + host mirror of the cells), CartesianDensityGrid (owner of one cmib_context) and the
 * FractalDensityMask.
 * Part of the host layer described in IonizationSimulation.hpp (class map, reference citations).
 */
#pragma once
#include "HostCommon.hpp"
#include "DensityFunctions.hpp"

namespace cmi {

/* ---- CartesianDensityGrid: host mirror of the cells + owner of the device context ---- */
/* geometry + host mirror of the cells: everything of the grid that needs no device (the
 * DensityFunction / DensityMask stage of IonizationSimulation::initialize) */
class CartesianCells {
public:
  CartesianCells(const SimulationBox &box, const std::array<int32_t, 3> &ncell)
      : anchor_(box.anchor), sides_(box.sides), ncell_(ncell), periodicity_(box.periodicity) {
    for (int k = 0; k < 3; ++k) cellside_[k] = sides_[k] / ncell_[k]; /* CartesianDensityGrid.cpp:80-86 */
    const size_t n = get_number_of_cells();
    number_density.assign(n, 0.);
    temperature.assign(n, 0.);
    ionic_fraction.assign(n * CMIB_NUM_IONS, 0.);
  }
  size_t get_number_of_cells() const { return (size_t)ncell_[0] * ncell_[1] * ncell_[2]; }
  /* long index ix*ny*nz + iy*nz + iz (CartesianDensityGrid.hpp:137-144) */
  Vec3 get_cell_midpoint(size_t index) const {
    const size_t nyz = (size_t)ncell_[1] * ncell_[2];
    const size_t ix = index / nyz, iy = (index % nyz) / ncell_[2], iz = index % ncell_[2];
    const size_t i[3] = {ix, iy, iz};
    Vec3 m;
    for (int k = 0; k < 3; ++k) m[k] = anchor_[k] + cellside_[k] * (double)i[k] + 0.5 * cellside_[k];
    return m;
  }
  double get_cell_volume() const { return cellside_[0] * cellside_[1] * cellside_[2]; }
  const std::array<int32_t, 3> &get_number_of_cells_3d() const { return ncell_; }
  const Vec3 &get_box_anchor() const { return anchor_; }
  const Vec3 &get_box_sides() const { return sides_; }
  /* DensityGrid::set_densities: evaluate the DensityFunction at every cell midpoint */
  void set_densities(DensityFunction &function) {
    if (function.set_densities(*this)) return;
    const size_t n = get_number_of_cells();
    for (size_t i = 0; i < n; ++i) {
      const DensityValues v = function(get_cell_midpoint(i));
      number_density[i] = v.number_density;
      temperature[i] = v.temperature;
      for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) ionic_fraction[(size_t)ion * n + i] = v.ionic_fraction[ion];
    }
  }

  /* host mirror, [ncell] and [14][ncell] in the reference's cell / ion order */
  std::vector<double> number_density, temperature, ionic_fraction;

protected:
  Vec3 anchor_, sides_, cellside_;
  std::array<int32_t, 3> ncell_;
  std::array<bool, 3> periodicity_;
};

class CartesianDensityGrid : public CartesianCells {
public:
  CartesianDensityGrid(const SimulationBox &box, const std::array<int32_t, 3> &ncell, int device = 0)
      : CartesianCells(box, ncell) {
    cmib_grid_desc d;
    for (int k = 0; k < 3; ++k) {
      d.anchor[k] = anchor_[k];
      d.sides[k] = sides_[k];
      d.ncell[k] = ncell_[k];
      d.periodic[k] = periodicity_[k] ? 1 : 0;
    }
    CMIB_CALL(cmib_create(&d, device, &ctx_));
  }
  CartesianDensityGrid(const SimulationBox &box, ParameterFile &params, int device = 0)
      : CartesianDensityGrid(box, params.get_value<std::array<int32_t, 3>>("DensityGrid:number of cells", {64, 64, 64}),
                             device) {}
  ~CartesianDensityGrid() {
    if (ctx_) cmib_destroy(ctx_);
  }
  CartesianDensityGrid(const CartesianDensityGrid &) = delete;
  CartesianDensityGrid &operator=(const CartesianDensityGrid &) = delete;

  /* DensityGrid::set_densities + upload */
  void initialize(DensityFunction &function) {
    set_densities(function);
    upload();
  }
  void upload() {
    CMIB_CALL(cmib_upload_cells(ctx_, number_density.data(), temperature.data(), ionic_fraction.data(), nullptr));
  }
  /* refresh the host mirror (for writers) */
  void download() {
    CMIB_CALL(cmib_download_cells(ctx_, number_density.data(), temperature.data(), ionic_fraction.data(), nullptr));
  }
  void reset_grid() { CMIB_CALL(cmib_reset_accumulators(ctx_)); }
  cmib_context *context() { return ctx_; }

private:
  cmib_context *ctx_ = nullptr;
};

inline bool GadgetSnapshotDensityFunction::set_densities(CartesianCells &grid) {
  const size_t ncell = grid.get_number_of_cells();
  const std::array<int32_t, 3> &nc = grid.get_number_of_cells_3d();
  const Vec3 &anchor = grid.get_box_anchor(), &sides = grid.get_box_sides();
  const double cs[3] = {sides[0] / nc[0], sides[1] / nc[1], sides[2] / nc[2]};
  const bool with_x = !neutral_fractions_.empty();
  std::vector<double> density(ncell, 0.), temperature(ncell, 0.), neutral(with_x ? ncell : 0, 0.);
  const size_t n = masses_.size();
  const int kmax = periodic_ ? 1 : 0;
  for (size_t i = 0; i < n; ++i) {
    const double h = smoothing_lengths_[i], m = masses_[i];
    const double p[3] = {positions_[3 * i], positions_[3 * i + 1], positions_[3 * i + 2]};
    for (int kx = -kmax; kx <= kmax; ++kx)
      for (int ky = -kmax; ky <= kmax; ++ky)
        for (int kz = -kmax; kz <= kmax; ++kz) {
          const int k[3] = {kx, ky, kz};
          /* cells whose midpoints anchor + (j + 1/2) cs can lie within h of this image of the particle */
          long lo[3], hi[3];
          bool empty = false;
          for (int d = 0; d < 3; ++d) {
            const double q = p[d] + k[d] * sides_[d];
            lo[d] = (long)std::ceil((q - h - anchor[d]) / cs[d] - 0.5) - 1; /* one cell of slack against rounding */
            hi[d] = (long)std::floor((q + h - anchor[d]) / cs[d] - 0.5) + 1;
            lo[d] = std::max(lo[d], 0l);
            hi[d] = std::min(hi[d], (long)nc[d] - 1);
            empty = empty || lo[d] > hi[d];
          }
          if (empty) continue;
          for (long ix = lo[0]; ix <= hi[0]; ++ix)
            for (long iy = lo[1]; iy <= hi[1]; ++iy)
              for (long iz = lo[2]; iz <= hi[2]; ++iz) {
                const size_t cell = ((size_t)ix * nc[1] + iy) * nc[2] + iz;
                const Vec3 x = grid.get_cell_midpoint(cell);
                double c[3];
                bool nearest = true;
                for (int d = 0; d < 3; ++d) {
                  c[d] = x[d] - p[d];
                  if (periodic_) { /* Box::periodic_distance; the pair counts for the image it picks */
                    const double c0 = c[d];
                    if (2 * c[d] < -sides_[d]) c[d] += sides_[d];
                    if (2 * c[d] >= sides_[d]) c[d] -= sides_[d];
                    nearest = nearest && (int)std::lround((c0 - c[d]) / sides_[d]) == k[d];
                  }
                }
                if (!nearest) continue;
                const double r = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
                const double u = r / h;
                if (!(u < 1.)) continue;
                const double splineval = m * kernel_evaluate(u, h);
                density[cell] += splineval;
                temperature[cell] += splineval * temperatures_[i] / densities_[i];
                if (with_x) neutral[cell] += splineval * neutral_fractions_[i];
              }
        }
  }
  for (size_t cidx = 0; cidx < ncell; ++cidx) {
    grid.number_density[cidx] = density[cidx] / 1.6737236e-27;
    grid.temperature[cidx] = temperature[cidx];
    for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) grid.ionic_fraction[(size_t)ion * ncell + cidx] = 0.;
    grid.ionic_fraction[cidx] = with_x ? neutral[cidx] / density[cidx] : 1.e-6;
    grid.ionic_fraction[ncell + cidx] = 1.e-6;
  }
  return true;
}

inline bool FLASHSnapshotDensityFunction::set_densities(CartesianCells &grid) {
  const size_t ncell = grid.get_number_of_cells();
  const std::array<int32_t, 3> &nc = grid.get_number_of_cells_3d();
  const Vec3 &anchor = grid.get_box_anchor(), &sides = grid.get_box_sides();
  const double cs[3] = {sides[0] / nc[0], sides[1] / nc[1], sides[2] / nc[2]};
  std::vector<char> filled(ncell, 0);
  for (const Block &b : blocks_) {
    long lo[3], hi[3];
    bool empty = false;
    for (int d = 0; d < 3; ++d) { /* midpoints anchor + (j + 1/2) cs inside [b.anchor, b.anchor + b.sides), one cell of slack */
      lo[d] = std::max((long)std::ceil((b.anchor[d] - anchor[d]) / cs[d] - 0.5) - 1, 0l);
      hi[d] = std::min((long)std::floor((b.anchor[d] + b.sides[d] - anchor[d]) / cs[d] - 0.5) + 1, (long)nc[d] - 1);
      empty = empty || lo[d] > hi[d];
    }
    if (empty) continue;
    for (long ix = lo[0]; ix <= hi[0]; ++ix)
      for (long iy = lo[1]; iy <= hi[1]; ++iy)
        for (long iz = lo[2]; iz <= hi[2]; ++iz) {
          const size_t cell = ((size_t)ix * nc[1] + iy) * nc[2] + iz;
          if (filled[cell]) continue; /* operator() takes the first block that contains the position */
          const Vec3 x = grid.get_cell_midpoint(cell);
          size_t c[3];
          bool inside = true;
          for (int d = 0; d < 3 && inside; ++d) {
            const double f = (x[d] - b.anchor[d]) / b.sides[d];
            inside = f >= 0. && f < 1.;
            if (inside) c[d] = std::min((size_t)(f * ncell_[d]), (size_t)ncell_[d] - 1);
          }
          if (!inside) continue;
          const size_t at = ((b.index * ncell_[2] + c[2]) * ncell_[1] + c[1]) * ncell_[0] + c[0];
          grid.number_density[cell] = densities_[at] / 1.6737236e-27;
          grid.temperature[cell] = (temperature_ <= 0.) ? temperatures_[at] : temperature_;
          for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) grid.ionic_fraction[(size_t)ion * ncell + cell] = 0.;
          grid.ionic_fraction[cell] = 1.e-6;
          grid.ionic_fraction[ncell + cell] = 1.e-6;
          filled[cell] = 1;
        }
  }
  for (size_t cell = 0; cell < ncell; ++cell)
    if (!filled[cell]) {
      const Vec3 x = grid.get_cell_midpoint(cell);
      cmi_error("Position [%g m, %g m, %g m] lies outside the blocks of snapshot \"%s\"!", x[0], x[1], x[2], filename_.c_str());
    }
  return true;
}

/* ---- DensityMask ---- */
/*
 * FractalDensityMask (src/FractalDensityMask.hpp:60-470, Elmegreen 1997): N^levels points placed by a
 * recursive random displacement (N points per level, length scale L = 10^(log10 N / D)) are counted
 * on a mask grid; apply() redistributes the gas of the cells inside the mask box in proportion
 * to the counts, keeping the total number of atoms.  Every first-level point owns a seed drawn
 * from RandomGenerator(seed), so the structure does not depend on threads; as in the reference
 * the job hand-out skips first-level index 0 (get_job increments before it reads, :246-253), i.e.
 * N - 1 of the N first-level points are generated.  Counts are integers and the sums of apply()
 * run in cell order: the masked grid is the reference's bit for bit (tests/test_host_layer.py).
 */
class FractalDensityMask {
public:
  FractalDensityMask(const Vec3 &box_anchor, const Vec3 &box_sides, const std::array<uint32_t, 3> &resolution,
                     uint32_t numpart, int32_t seed, double fractal_dimension, uint32_t num_level, double fractal_fraction)
      : anchor_(box_anchor), sides_(box_sides), resolution_(resolution),
        N_((uint32_t)std::ceil(std::pow(numpart, 1. / num_level))),
        L_(std::pow(10., std::log10(N_) / fractal_dimension)), num_level_(num_level),
        fractal_fraction_(fractal_fraction),
        distribution_((size_t)resolution[0] * resolution[1] * resolution[2], 0) {
    first_level_seeds_.resize(N_, 0);
    RandomGenerator random_generator(seed);
    for (uint32_t i = 0; i < N_; ++i) first_level_seeds_[i] = random_generator.get_random_integer();
  }
  explicit FractalDensityMask(ParameterFile &params)
      : FractalDensityMask(
            params.get_physical_vector<QUANTITY_LENGTH>("DensityMask:box anchor", "[-5. pc, -5. pc, -5. pc]"),
            params.get_physical_vector<QUANTITY_LENGTH>("DensityMask:box sides", "[10. pc, 10. pc, 10. pc]"),
            params.get_value<std::array<uint32_t, 3>>("DensityMask:resolution", {20, 20, 20}),
            params.get_value<uint32_t>("DensityMask:number of particles", 1000000),
            params.get_value<int32_t>("DensityMask:random seed", 42),
            params.get_value<double>("DensityMask:fractal dimension", 2.6),
            params.get_value<uint32_t>("DensityMask:number of levels", 4),
            params.get_value<double>("DensityMask:fractal fraction", 1.)) {}

  void initialize() {
    for (uint32_t index = 1; index < N_; ++index) {
      RandomGenerator random_generator(first_level_seeds_[index]);
      make_fractal_grid(random_generator, {0., 0., 0.}, 1);
    }
  }

  /* number_density in the grid's cell order; midpoint(i) and the (uniform) cell volume of the grid */
  template <class Grid> void apply(Grid &grid) const {
    const double smooth_fraction = 1. - fractal_fraction_;
    const size_t n = grid.get_number_of_cells();
    const double volume = grid.get_cell_volume();
    double Ntot = 0., Nsmooth = 0., Nfractal = 0.;
    for (size_t i = 0; i < n; ++i) {
      const Vec3 midpoint = grid.get_cell_midpoint(i);
      if (!inside(midpoint)) continue;
      const double Ncell = grid.number_density[i] * volume;
      Ntot += Ncell;
      Nsmooth += smooth_fraction * Ncell;
      Nfractal += fractal_fraction_ * Ncell * distribution_[index(midpoint)];
    }
    const double fractal_norm = (Ntot - Nsmooth) / Nfractal;
    for (size_t i = 0; i < n; ++i) {
      const Vec3 midpoint = grid.get_cell_midpoint(i);
      if (!inside(midpoint)) continue;
      const double ncell = grid.number_density[i];
      const double nsmooth = smooth_fraction * ncell;
      const double nfractal = fractal_fraction_ * fractal_norm * ncell * distribution_[index(midpoint)];
      grid.number_density[i] = nsmooth + nfractal;
    }
  }
  const std::vector<uint64_t> &distribution() const { return distribution_; }

private:
  /* Box::inside (src/Box.hpp): anchor <= x < anchor + sides per coordinate */
  bool inside(const Vec3 &p) const {
    for (int d = 0; d < 3; ++d)
      if (!(p[d] >= anchor_[d] && p[d] < anchor_[d] + sides_[d])) return false;
    return true;
  }
  size_t index(const Vec3 &p) const {
    size_t idx[3];
    for (int d = 0; d < 3; ++d) idx[d] = (size_t)((p[d] - anchor_[d]) / sides_[d] * resolution_[d]);
    return (idx[0] * resolution_[1] + idx[1]) * resolution_[2] + idx[2];
  }
  void make_fractal_grid(RandomGenerator &random_generator, Vec3 x_level, uint32_t current_level) {
    for (int d = 0; d < 3; ++d)
      x_level[d] += 2. * (random_generator.get_uniform_random_double() - 0.5) / std::pow(L_, current_level);
    if (current_level < num_level_) {
      for (uint32_t i = 0; i < N_; ++i) make_fractal_grid(random_generator, x_level, current_level + 1);
      return;
    }
    size_t idx[3];
    for (int d = 0; d < 3; ++d) {
      x_level[d] *= 0.5 * L_;
      x_level[d] += 0.5;
      if (x_level[d] < 0.) x_level[d] += 1.;
      if (x_level[d] >= 1.) x_level[d] -= 1.;
      idx[d] = (size_t)(x_level[d] * resolution_[d]);
    }
    ++distribution_[(idx[0] * resolution_[1] + idx[1]) * resolution_[2] + idx[2]];
  }
  Vec3 anchor_, sides_;
  std::array<uint32_t, 3> resolution_;
  uint32_t N_;
  double L_;
  uint32_t num_level_;
  double fractal_fraction_;
  std::vector<int32_t> first_level_seeds_;
  std::vector<uint64_t> distribution_;
};

struct DensityMaskFactory {
  static FractalDensityMask *generate(ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("DensityMask:type", "None");
    if (log) log->write_info("Requested DensityMask type: ", type);
    if (type == "Fractal") return new FractalDensityMask(params);
    if (type == "None") return nullptr;
    cmi_error("Unknown DensityMask type: \"%s\"!", type.c_str());
  }
};

} // namespace cmi
