/*
 * DensityFunctions.hpp — DensityFunction plugins of the host layer (DensityFunctionFactory.hpp): closed forms, tables, snapshots.
 * Evaluated once per run on the host to fill the initial grid; nothing of the iteration.
 * Part of the host layer described in IonizationSimulation.hpp (class map, reference citations).
 */
#pragma once
#include "HostCommon.hpp"

namespace cmi {

/* ---- DensityFunction ---- */
struct DensityValues {
  double number_density = 0.;
  double temperature = 0.;
  double ionic_fraction[CMIB_NUM_IONS] = {0.};
  double cosmic_ray_factor = -1.; /* DensityValues.hpp:65-71 */
};

class CartesianCells;
class DensityFunction {
public:
  virtual ~DensityFunction() {}
  virtual void initialize() {}
  virtual DensityValues operator()(const Vec3 &cell_midpoint) = 0;
  /* a function that fills the whole grid at once (SPHArrayInterface) returns true here */
  virtual bool set_densities(CartesianCells &) { return false; }
};

class HomogeneousDensityFunction : public DensityFunction {
public:
  HomogeneousDensityFunction(double density, double temperature, double neutral_fraction_H)
      : density_(density), temperature_(temperature), neutral_fraction_H_(neutral_fraction_H) {}
  explicit HomogeneousDensityFunction(ParameterFile &params)
      : HomogeneousDensityFunction(
            params.get_physical_value<QUANTITY_NUMBER_DENSITY>("DensityFunction:density", "100. cm^-3"),
            params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "8000. K"),
            params.get_value<double>("DensityFunction:neutral fraction H", 1.e-6)) {}
  DensityValues operator()(const Vec3 &) override {
    DensityValues v;
    v.number_density = density_;
    v.temperature = temperature_;
    v.ionic_fraction[0] = neutral_fraction_H_;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }

private:
  double density_, temperature_, neutral_fraction_H_;
};

class BlockSyntaxDensityFunction : public DensityFunction {
  struct Block {
    Vec3 origin, sides;
    double exponent, number_density, temperature, neutral_fraction_H;
    bool is_inside(const Vec3 &p) const {
      double r = 0.;
      for (int i = 0; i < 3; ++i) {
        const double x = 2. * std::abs(p[i] - origin[i]) / sides[i];
        if (exponent < 10.) r += std::pow(x, exponent);
        else r = std::max(r, x);
      }
      if (exponent < 10.) r = std::pow(r, 1. / exponent);
      return r <= 1.;
    }
  };

public:
  explicit BlockSyntaxDensityFunction(const std::string &filename) {
    std::ifstream file(filename);
    if (!file) cmi_error("Error while opening file \"%s\"!", filename.c_str());
    YAMLDictionary blockfile(file);
    const uint32_t numblock = blockfile.get_value<uint32_t>("number of blocks");
    for (uint32_t i = 0; i < numblock; ++i) {
      const std::string name = "block[" + std::to_string(i) + "]:";
      Block b;
      b.origin = blockfile.get_physical_vector<QUANTITY_LENGTH>(name + "origin");
      b.sides = blockfile.get_physical_vector<QUANTITY_LENGTH>(name + "sides");
      const std::string type = blockfile.get_value<std::string>(name + "type");
      if (type == "rhombus") b.exponent = 1.;
      else if (type == "sphere") b.exponent = 2.;
      else if (type == "cube") b.exponent = 10.;
      else cmi_error("Unknown block type: \"%s\"!", type.c_str());
      if (blockfile.has_value(name + "number density")) {
        b.number_density = blockfile.get_physical_value<QUANTITY_NUMBER_DENSITY>(name + "number density");
      } else {
        b.number_density = blockfile.get_physical_value<QUANTITY_DENSITY>(name + "density");
        b.number_density /= constants::proton_mass;
      }
      b.temperature = blockfile.get_physical_value<QUANTITY_TEMPERATURE>(name + "initial temperature");
      b.neutral_fraction_H = blockfile.get_value<double>(name + "neutral fraction H", 1.e-6);
      (void)blockfile.get_physical_vector<QUANTITY_VELOCITY>(name + "initial velocity", "[0. m s^-1, 0. m s^-1, 0. m s^-1]");
      if (b.number_density < 0.) cmi_error("Negative density (%g) given for block %u!", b.number_density, i);
      if (b.temperature < 0.) cmi_error("Negative temperature (%g) given for block %u!", b.temperature, i);
      blocks_.push_back(b);
    }
    std::ofstream ofile(filename + ".used-values");
    blockfile.print_contents(ofile, true);
  }
  explicit BlockSyntaxDensityFunction(ParameterFile &params)
      : BlockSyntaxDensityFunction(params.get_filename("DensityFunction:filename")) {}

  DensityValues operator()(const Vec3 &position) override {
    double density = -1., temperature = -1., xH = -1.;
    for (const Block &b : blocks_) { /* later blocks win */
      if (b.is_inside(position)) {
        density = b.number_density;
        temperature = b.temperature;
        xH = b.neutral_fraction_H;
      }
    }
    if (density < 0. || temperature < 0. || xH < 0.)
      cmi_error("No block found containing position [%g m, %g m, %g m]!", position[0], position[1], position[2]);
    DensityValues v;
    v.number_density = density;
    v.temperature = temperature;
    v.ionic_fraction[0] = xH;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }

private:
  std::vector<Block> blocks_;
};

/* AsciiFileDensityFunction (src/AsciiFileDensityFunction.cpp:40-186): "x y z density" rows on a regular
 * grid of its own (not necessarily the simulation grid); a cell takes the value of the file cell
 * its midpoint falls in */
class AsciiFileDensityFunction : public DensityFunction {
public:
  AsciiFileDensityFunction(const std::string &filename, const std::array<uint32_t, 3> &ncell, const Vec3 &anchor,
                           const Vec3 &sides, double temperature, double length_unit_in_SI, double density_unit_in_SI)
      : ncell_(ncell), anchor_(anchor), sides_(sides), temperature_(temperature),
        grid_((size_t)ncell[0] * ncell[1] * ncell[2], -1.) {
    std::ifstream file(filename);
    if (!file.is_open()) cmi_error("Could not open file \"%s\"!", filename.c_str());
    std::string line;
    while (getline(file, line)) {
      if (line[0] == '#') continue;
      double x = 0., y = 0., z = 0., rho = 0.;
      std::stringstream linestream(line);
      linestream >> x >> y >> z >> rho;
      x *= length_unit_in_SI;
      y *= length_unit_in_SI;
      z *= length_unit_in_SI;
      rho *= density_unit_in_SI;
      grid_[index({x, y, z})] = rho;
    }
    for (uint32_t i = 0; i < ncell_[0]; ++i)
      for (uint32_t j = 0; j < ncell_[1]; ++j)
        for (uint32_t k = 0; k < ncell_[2]; ++k)
          if (grid_[((size_t)i * ncell_[1] + j) * ncell_[2] + k] < 0.)
            cmi_error("No value found for cell [%u, %u, %u]!", i, j, k);
  }
  explicit AsciiFileDensityFunction(ParameterFile &params)
      : AsciiFileDensityFunction(
            params.get_filename("DensityFunction:filename"),
            params.get_value<std::array<uint32_t, 3>>("DensityFunction:number of cells", {64, 64, 64}),
            params.get_physical_vector<QUANTITY_LENGTH>("DensityFunction:box anchor", "[-5. pc, -5. pc, -5. pc]"),
            params.get_physical_vector<QUANTITY_LENGTH>("DensityFunction:box sides", "[10. pc, 10. pc, 10. pc]"),
            params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "8000. K"),
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:length unit", "1. m"),
            params.get_physical_value<QUANTITY_NUMBER_DENSITY>("DensityFunction:density unit", "1. m^-3")) {}

  DensityValues operator()(const Vec3 &position) override {
    DensityValues v;
    v.number_density = grid_[index(position)];
    v.temperature = temperature_;
    v.ionic_fraction[0] = 1.e-6;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }

private:
  /* (p - anchor) / sides * ncell, truncated (.cpp:82-85, 170-176); out-of-range rows are the
   * reference's undefined behaviour: here an error */
  size_t index(const Vec3 &p) const {
    size_t idx[3];
    for (int d = 0; d < 3; ++d) {
      const double f = (p[d] - anchor_[d]) / sides_[d] * ncell_[d];
      if (!(f >= 0.) || !(f < (double)ncell_[d]))
        cmi_error("Position [%g m, %g m, %g m] outside the box of the AsciiFile density grid!", p[0], p[1], p[2]);
      idx[d] = (size_t)f;
    }
    return (idx[0] * ncell_[1] + idx[1]) * ncell_[2] + idx[2];
  }
  std::array<uint32_t, 3> ncell_;
  Vec3 anchor_, sides_;
  double temperature_;
  std::vector<double> grid_;
};

/* InterpolatedDensityFunction (src/InterpolatedDensityFunction.cpp:40-369): a 1-, 2- or 3-D table of
 * number densities (a YAML header between two "---" lines names the columns and their units, rows
 * follow with x slowest / z fastest), trilinear interpolation at the cell midpoint; an axis with
 * fewer than two points is constant between its bounds.  Like the reference's reader this one
 * never rewinds an axis index while reading rows (:213-247), i.e. tables with ONE non-trivial axis
 * are what works; where the reference then writes out of bounds this reader reports an error. */
class InterpolatedDensityFunction : public DensityFunction {
public:
  InterpolatedDensityFunction(const std::string &filename, double temperature) : temperature_(temperature) {
    std::ifstream file(filename);
    if (!file) cmi_error("Error while opening file \"%s\"!", filename.c_str());
    std::string line;
    while (std::getline(file, line) && line != "---") {
    }
    if (line != "---") cmi_error("No YAML block found in file \"%s\"!", filename.c_str());
    std::string yaml_block;
    while (std::getline(file, line) && line != "---") yaml_block += line + "\n";
    if (line != "---") cmi_error("Reached end of file \"%s\" while parsing YAML block!", filename.c_str());
    std::istringstream yaml_stream(yaml_block);
    YAMLDictionary yaml(yaml_stream);
    const char *axis_name[3] = {"x", "y", "z"};
    uint32_t num[3];
    for (int d = 0; d < 3; ++d) num[d] = yaml.get_value<uint32_t>(std::string("num_") + axis_name[d]);
    for (int d = 0; d < 3; ++d) {
      bounds_[d][0] = yaml.get_physical_value<QUANTITY_LENGTH>(std::string(axis_name[d]) + "min");
      bounds_[d][1] = yaml.get_physical_value<QUANTITY_LENGTH>(std::string(axis_name[d]) + "max");
    }
    const uint32_t num_column = yaml.get_value<uint32_t>("num_column");
    std::map<std::string, uint32_t> name_to_column;
    std::vector<std::string> units(num_column);
    for (uint32_t i = 0; i < num_column; ++i) {
      const std::string column = "column_" + std::to_string(i) + "_";
      const std::string name = yaml.get_value<std::string>(column + "variable");
      units[i] = yaml.get_value<std::string>(column + "unit");
      name_to_column[name] = i;
    }
    if (num[0] == 0 && num[1] == 0 && num[2] == 0)
      cmi_error("No coordinate values provided! We need at least one non-trivial coordinate axis.");
    const char *axis_upper[3] = {"X", "Y", "Z"};
    for (int d = 0; d < 3; ++d)
      if (bounds_[d][0] > bounds_[d][1]) cmi_error("Minimal %s value larger than maximal %s value!", axis_upper[d], axis_upper[d]);
    uint32_t column_of[3] = {0, 0, 0};
    for (int d = 0; d < 3; ++d) {
      if (num[d] != 0) {
        if (name_to_column.count(axis_name[d]) == 0) cmi_error("No column found containing %s values!", axis_name[d]);
        column_of[d] = name_to_column[axis_name[d]];
      }
      if (num[d] > 1) {
        coords_[d].assign(num[d], 0.);
      } else {
        coords_[d] = {bounds_[d][0], bounds_[d][1]};
      }
    }
    if (name_to_column.count("number density") == 0) cmi_error("No column found containing number density values!");
    const uint32_t density_column = name_to_column["number density"];
    const size_t ny = coords_[1].size(), nz = coords_[2].size();
    densities_.assign(coords_[0].size() * ny * nz, 0.);
    size_t idx[3] = {0, 0, 0}, i = 0;
    while (std::getline(file, line)) {
      std::stringstream lstream(line);
      std::vector<double> row(num_column);
      for (uint32_t j = 0; j < num_column; ++j) lstream >> row[j];
      for (int d = 0; d < 3; ++d) {
        if (num[d] == 0) continue;
        const double next = UnitConverter::to_SI(QUANTITY_LENGTH, row[column_of[d]], units[column_of[d]]);
        if (i > 0 && next != coords_[d][idx[d]]) {
          ++idx[d];
          if (idx[d] >= coords_[d].size())
            cmi_error("Too many different %s values in file \"%s\"!", axis_name[d], filename.c_str());
        }
        coords_[d][idx[d]] = next;
      }
      densities_[(idx[0] * ny + idx[1]) * nz + idx[2]] =
          UnitConverter::to_SI(QUANTITY_NUMBER_DENSITY, row[density_column], units[density_column]);
      ++i;
    }
    /* complete the axes that have a single value (:264-289) */
    const size_t nx = coords_[0].size();
    if (num[0] < 2)
      for (size_t iy = 0; iy < ny; ++iy)
        for (size_t iz = 0; iz < nz; ++iz) densities_[(1 * ny + iy) * nz + iz] = densities_[(0 * ny + iy) * nz + iz];
    if (num[1] < 2)
      for (size_t ix = 0; ix < nx; ++ix)
        for (size_t iz = 0; iz < nz; ++iz) densities_[(ix * ny + 1) * nz + iz] = densities_[(ix * ny + 0) * nz + iz];
    if (num[2] < 2)
      for (size_t ix = 0; ix < nx; ++ix)
        for (size_t iy = 0; iy < ny; ++iy) densities_[(ix * ny + iy) * nz + 1] = densities_[(ix * ny + iy) * nz + 0];
  }
  explicit InterpolatedDensityFunction(ParameterFile &params)
      : InterpolatedDensityFunction(params.get_filename("DensityFunction:filename"),
                                    params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "8000. K")) {}

  DensityValues operator()(const Vec3 &position) override {
    size_t i[3];
    double w[3], omw[3];
    for (int d = 0; d < 3; ++d) {
      i[d] = locate_bin(position[d], coords_[d].data(), (uint32_t)coords_[d].size());
      w[d] = (position[d] - coords_[d][i[d]]) / (coords_[d][i[d] + 1] - coords_[d][i[d]]);
      omw[d] = 1. - w[d];
    }
    const size_t ny = coords_[1].size(), nz = coords_[2].size();
    auto n = [&](size_t ix, size_t iy, size_t iz) { return densities_[(ix * ny + iy) * nz + iz]; };
    const double c00 = n(i[0], i[1], i[2]) * omw[0] + n(i[0] + 1, i[1], i[2]) * w[0];
    const double c01 = n(i[0], i[1], i[2] + 1) * omw[0] + n(i[0] + 1, i[1], i[2] + 1) * w[0];
    const double c10 = n(i[0], i[1] + 1, i[2]) * omw[0] + n(i[0] + 1, i[1] + 1, i[2]) * w[0];
    const double c11 = n(i[0], i[1] + 1, i[2] + 1) * omw[0] + n(i[0] + 1, i[1] + 1, i[2] + 1) * w[0];
    const double c0 = c00 * omw[1] + c10 * w[1];
    const double c1 = c01 * omw[1] + c11 * w[1];
    DensityValues v;
    v.number_density = c0 * omw[2] + c1 * w[2];
    v.temperature = temperature_;
    v.ionic_fraction[0] = 1.e-6;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }

private:
  /* Utilities::locate (src/Utilities.hpp:726-742) */
  static size_t locate_bin(double x, const double *xarr, uint32_t length) {
    uint32_t jl = 0, ju = length;
    while (ju - jl > 1) {
      const uint32_t jm = (ju + jl) >> 1;
      if (x > xarr[jm]) jl = jm; else ju = jm;
    }
    if (jl == length - 1) --jl;
    return jl;
  }
  double temperature_;
  double bounds_[3][2];
  std::vector<double> coords_[3];
  std::vector<double> densities_;
};

/* ---- analytic density profiles (each a closed form per cell midpoint; operation order of the
 * reference, so that the initial grid is the same doubles) ---- */

/* isothermal gas in the potential of a cored dark-matter halo
 * (CoredDMProfileDensityFunction.hpp:84-156) */
class CoredDMProfileDensityFunction : public DensityFunction {
public:
  CoredDMProfileDensityFunction(double r0, double vinf, double rho0, double temperature, double neutral_fraction,
                                double gamma = 1.)
      : r0inv_(1. / r0), vratio_(gamma * vinf * vinf / sound_speed_squared(neutral_fraction, temperature)),
        n0_(rho0 / mean_particle_mass(neutral_fraction)), temperature_(temperature / gamma),
        neutral_fraction_(neutral_fraction) {}
  explicit CoredDMProfileDensityFunction(ParameterFile &params)
      : CoredDMProfileDensityFunction(
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:core radius", "300. pc"),
            params.get_physical_value<QUANTITY_VELOCITY>("DensityFunction:maximum circular velocity", "21.1 km s^-1"),
            params.get_physical_value<QUANTITY_DENSITY>("DensityFunction:central density", "9.48e-21 g cm^-3"),
            params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "500. K"),
            params.get_value<double>("DensityFunction:neutral fraction", 1.),
            params.get_value<double>("DensityFunction:polytropic index", 1.)) {}
  DensityValues operator()(const Vec3 &x) override {
    const double r = std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double ksi = r * r0inv_;
    DensityValues v;
    v.number_density = n0_ * std::exp(-vratio_ * (0.5 * std::log(1. + ksi * ksi) + std::atan(ksi) / ksi - 1.));
    v.temperature = temperature_;
    v.ionic_fraction[0] = neutral_fraction_;
    return v;
  }

private:
  static double mean_particle_mass(double neutral_fraction) {
    return 0.5 * constants::proton_mass * (1. + neutral_fraction);
  }
  static double sound_speed_squared(double neutral_fraction, double temperature) {
    return constants::boltzmann * temperature / mean_particle_mass(neutral_fraction);
  }
  double r0inv_, vratio_, n0_, temperature_, neutral_fraction_;
};

/* power-law envelope around a point mass, scaled by its Bondi radius (DiscICDensityFunction.hpp:123-186);
 * the rotation velocity of that profile belongs to the hydro and is not part of the grid here */
class DiscICDensityFunction : public DensityFunction {
public:
  DiscICDensityFunction(double mass, double temperature, double rho_B, double gamma_rho)
      : R_B_(0.5 * constants::newton_constant * mass * mean_particle_mass(temperature) /
             (constants::boltzmann * temperature)),
        n_B_(rho_B / mean_particle_mass(temperature)), gamma_rho_(gamma_rho), temperature_(temperature),
        neutral_fraction_H_(temperature < 1.e4 ? 1. : 1.e-6) {}
  explicit DiscICDensityFunction(ParameterFile &params)
      : DiscICDensityFunction(params.get_physical_value<QUANTITY_MASS>("DensityFunction:mass", "20. Msol"),
                              params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "500. K"),
                              params.get_physical_value<QUANTITY_DENSITY>("DensityFunction:Bondi density", "3.1e3 g m^-3"),
                              params.get_value<double>("DensityFunction:density power", 1.5)) {
    params.get_physical_value<QUANTITY_VELOCITY>("DensityFunction:Bondi velocity", "2.873 km s^-1");
    params.get_value<double>("DensityFunction:velocity power", 0.5);
  }
  DensityValues operator()(const Vec3 &x) override {
    const double rinv = R_B_ / std::sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    DensityValues v;
    v.number_density = n_B_ * std::pow(rinv, gamma_rho_);
    v.temperature = temperature_;
    v.ionic_fraction[0] = neutral_fraction_H_;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }

private:
  static double mean_particle_mass(double temperature) {
    return temperature < 1.e4 ? constants::proton_mass : 0.5 * constants::proton_mass;
  }
  double R_B_, n_B_, gamma_rho_, temperature_, neutral_fraction_H_;
};

/* vertical gas profile of a patch of a galactic disc in equilibrium with a stellar sech^2 disc
 * (DiscPatchDensityFunction.hpp:120-176) */
class DiscPatchDensityFunction : public DensityFunction {
public:
  DiscPatchDensityFunction(double disc_z, double surface_density, double scale_height, double gas_fraction,
                           double temperature, double neutral_fraction)
      : disc_z_(disc_z), b_inv_(1. / scale_height),
        exponent_(-2. * scale_height / gas_disc_scale_height(surface_density, temperature, neutral_fraction)),
        density_norm_(0.5 * gas_fraction * surface_density * mass_fraction_factor(exponent_) * b_inv_ /
                      constants::proton_mass),
        temperature_(temperature), neutral_fraction_(neutral_fraction) {}
  explicit DiscPatchDensityFunction(ParameterFile &params)
      : DiscPatchDensityFunction(
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:disc z", "0. m"),
            params.get_physical_value<QUANTITY_SURFACE_DENSITY>("DensityFunction:surface density", "30. Msol pc^-2"),
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:scale height", "200. pc"),
            params.get_value<double>("DensityFunction:gas fraction", 0.1),
            params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "1.e4 K"),
            params.get_value<double>("DensityFunction:neutral fraction", 1e-6)) {}
  DensityValues operator()(const Vec3 &x) override {
    const double dz = x[2] - disc_z_;
    DensityValues v;
    v.number_density = density_norm_ * std::pow(std::cosh(dz * b_inv_), exponent_);
    v.temperature = temperature_;
    v.ionic_fraction[0] = neutral_fraction_;
    return v;
  }

private:
  static double gas_disc_scale_height(double surface_density, double temperature, double neutral_fraction) {
    return (constants::boltzmann * temperature) /
           (0.5 * constants::proton_mass * (1. + neutral_fraction) * M_PI * constants::newton_constant * surface_density);
  }
  /* the reference's cubic fit (in log10) of the mass integral of cosh^exponent */
  static double mass_fraction_factor(double exponent) {
    const double x = std::log10(-0.5 * exponent);
    const double x2 = x * x;
    const double y = 0.01499337 * x2 * x - 0.08454788 * x2 + 0.63503798 * x - 0.01018254;
    return std::pow(10., y);
  }
  double disc_z_, b_inv_, exponent_, density_norm_, temperature_, neutral_fraction_;
};

/* double-exponential disc of a spiral galaxy, cut at 15 kpc (SpiralGalaxyDensityFunction.hpp:69-131).
 * As in the reference the central *number* density is multiplied by 1.674e-27 (a hydrogen mass in kg)
 * before it is stored as the cells' number density, the gas is neutral and the temperature is 0. */
class SpiralGalaxyDensityFunction : public DensityFunction {
public:
  SpiralGalaxyDensityFunction(double r_ISM, double h_ISM, double n_0)
      : r_ISM_(r_ISM), h_ISM_(h_ISM), n_0_(1.674e-27 * n_0), kpc_(3.086e19) {}
  explicit SpiralGalaxyDensityFunction(ParameterFile &params)
      : SpiralGalaxyDensityFunction(
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:scale length ISM", "6. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:scale height ISM", "0.22 kpc"),
            params.get_physical_value<QUANTITY_NUMBER_DENSITY>("DensityFunction:central density", "1. cm^-3")) {}
  DensityValues operator()(const Vec3 &x) override {
    const double w = std::sqrt(x[0] * x[0] + x[1] * x[1]);
    DensityValues v;
    if (w < 15. * kpc_ && std::abs(x[2]) < 15. * kpc_)
      v.number_density = n_0_ * std::exp(-w / r_ISM_) * std::exp(-std::abs(x[2]) / h_ISM_);
    v.temperature = 0.;
    v.ionic_fraction[0] = 1.;
    v.ionic_fraction[1] = 0.;
    return v;
  }

private:
  double r_ISM_, h_ISM_, n_0_, kpc_;
};

/* A snapshot of an earlier run as initial condition (CMacIonizeSnapshotDensityFunction.cpp:108-470, :504-523):
 * reads /Parameters (box, number of cells, grid type), /Units and /PartType0/{Coordinates, NumberDensity,
 * Temperature, NeutralFraction<ion>} of a Gadget-style snapshot written by the reference or by this host layer
 * (host/HDF5Reader.hpp, no HDF5 library) and returns, for a position, the values of the snapshot cell that
 * contains it.  Cartesian snapshots place a cell by its coordinates, task-based ones by the subgrid order of
 * the cells; snapshots of AMR / Voronoi grids are refused (those grids are outside the accelerated path), and
 * so are the hydro variants (`use density`, `use pressure`). */
class CMacIonizeSnapshotDensityFunction : public DensityFunction {
public:
  CMacIonizeSnapshotDensityFunction(std::string filename, bool use_density, bool use_pressure,
                                    double initial_neutral_fraction)
      : filename_(std::move(filename)), initial_neutral_fraction_(initial_neutral_fraction) {
    if (use_density || use_pressure)
      cmi_error("DensityFunction:use density / use pressure read hydro snapshots, which the B200 backend does not provide!");
  }
  explicit CMacIonizeSnapshotDensityFunction(ParameterFile &params)
      : CMacIonizeSnapshotDensityFunction(params.get_filename("DensityFunction:filename"),
                                          params.get_value<bool>("DensityFunction:use density", false),
                                          params.get_value<bool>("DensityFunction:use pressure", false),
                                          params.get_value<double>("DensityFunction:initial neutral fraction", 1.e-6)) {}
  void initialize() override {
    hdf5::HDF5Input file(filename_);
    YAMLDictionary parameters;
    for (const std::string &name : file.get_attribute_names("/Parameters"))
      parameters.add_value(name, file.read_string_attribute("/Parameters", name));
    anchor_ = parameters.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:anchor");
    sides_ = parameters.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:sides");
    ncell_ = parameters.get_value<std::array<uint32_t, 3>>("DensityGrid:number of cells");
    const std::string type = parameters.has_value("DensityGrid:type") ? parameters.get_value<std::string>("DensityGrid:type")
                                                                       : std::string("TaskBased");
    if (type != "Cartesian" && type != "TaskBased")
      cmi_error("Snapshot \"%s\" holds a %s grid; the B200 backend reads Cartesian and TaskBased snapshots!",
                filename_.c_str(), type.c_str());
    double unit_length_in_SI = 1., unit_density_in_SI = 1., unit_temperature_in_SI = 1.;
    if (file.exists("/Units")) {
      const double unit_length_in_cgs = file.read_double_attribute("/Units", "Unit length in cgs (U_L)")[0];
      unit_temperature_in_SI = file.read_double_attribute("/Units", "Unit temperature in cgs (U_T)")[0];
      unit_length_in_SI = UnitConverter::to_SI(QUANTITY_LENGTH, unit_length_in_cgs, "cm");
      unit_density_in_SI = 1. / unit_length_in_SI / unit_length_in_SI / unit_length_in_SI;
    }
    if (!file.exists("/PartType0/NumberDensity"))
      cmi_error("Snapshot \"%s\" holds no NumberDensity (hydro snapshots are not provided by the B200 backend)!", filename_.c_str());
    if (!file.exists("/PartType0/Temperature"))
      cmi_error("Snapshot \"%s\" holds no Temperature (switch on DensityGridWriterFields:Temperature in the run that writes it)!",
                filename_.c_str());
    std::vector<double> densities = file.read_dataset("/PartType0/NumberDensity");
    std::vector<double> temperatures = file.read_dataset("/PartType0/Temperature");
    const size_t n = densities.size();
    const size_t ntot = (size_t)ncell_[0] * ncell_[1] * ncell_[2];
    if (n != ntot || temperatures.size() != n)
      cmi_error("Snapshot \"%s\": %zu cells in /PartType0, %zu in /Parameters!", filename_.c_str(), n, ntot);
    std::vector<std::vector<double>> fractions(CMIB_NUM_IONS);
    for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) {
      const std::string name = std::string("/PartType0/NeutralFraction") + ion_symbol(ion);
      if (file.exists(name)) fractions[ion] = file.read_dataset(name);
      else fractions[ion].assign(n, initial_neutral_fraction_);
      if (fractions[ion].size() != n) cmi_error("Snapshot \"%s\": %s has the wrong size!", filename_.c_str(), name.c_str());
    }
    for (size_t i = 0; i < n; ++i) {
      densities[i] *= unit_density_in_SI;
      temperatures[i] *= unit_temperature_in_SI;
    }
    /* slot of snapshot cell i in the ix*ny*nz + iy*nz + iz order */
    std::vector<size_t> slot(n);
    if (type == "Cartesian") {
      std::vector<uint64_t> dims;
      std::vector<double> x = file.read_dataset("/PartType0/Coordinates", &dims);
      if (dims.size() != 2 || dims[0] != n || dims[1] != 3) cmi_error("Snapshot \"%s\": bad Coordinates!", filename_.c_str());
      for (size_t i = 0; i < n; ++i) {
        size_t idx[3];
        for (int k = 0; k < 3; ++k) {
          idx[k] = (size_t)(ncell_[k] * (x[3 * i + k] * unit_length_in_SI) / sides_[k]);
          if (idx[k] >= ncell_[k]) cmi_error("Snapshot \"%s\": cell %zu lies outside the box!", filename_.c_str(), i);
        }
        slot[i] = (idx[0] * ncell_[1] + idx[1]) * ncell_[2] + idx[2];
      }
    } else {
      const auto nsub = parameters.get_value<std::array<uint32_t, 3>>("DensitySubGridCreator:number of subgrids");
      const size_t nb[3] = {ncell_[0] / nsub[0], ncell_[1] / nsub[1], ncell_[2] / nsub[2]};
      const size_t nbtot = nb[0] * nb[1] * nb[2];
      for (size_t six = 0; six < nsub[0]; ++six)
        for (size_t siy = 0; siy < nsub[1]; ++siy)
          for (size_t siz = 0; siz < nsub[2]; ++siz) {
            const size_t subgrid = (six * nsub[1] + siy) * nsub[2] + siz;
            for (size_t cix = 0; cix < nb[0]; ++cix)
              for (size_t ciy = 0; ciy < nb[1]; ++ciy)
                for (size_t ciz = 0; ciz < nb[2]; ++ciz) {
                  const size_t cell = subgrid * nbtot + (cix * nb[1] + ciy) * nb[2] + ciz;
                  if (cell >= n) cmi_error("Snapshot \"%s\": subgrids do not match the number of cells!", filename_.c_str());
                  slot[cell] = ((six * nb[0] + cix) * ncell_[1] + (siy * nb[1] + ciy)) * ncell_[2] + (siz * nb[2] + ciz);
                }
          }
    }
    values_.assign(n, DensityValues());
    std::vector<char> filled(n, 0);
    for (size_t i = 0; i < n; ++i) {
      DensityValues &v = values_[slot[i]];
      v.number_density = densities[i];
      v.temperature = temperatures[i];
      for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) v.ionic_fraction[ion] = fractions[ion][i];
      filled[slot[i]] = 1;
    }
    for (size_t i = 0; i < n; ++i)
      if (!filled[i])
        cmi_error("No values found for cell (%zu, %zu, %zu)!", i / ((size_t)ncell_[1] * ncell_[2]),
                  (i / ncell_[2]) % ncell_[1], i % ncell_[2]);
  }
  DensityValues operator()(const Vec3 &x) override {
    size_t idx[3];
    for (int k = 0; k < 3; ++k) {
      idx[k] = (size_t)(ncell_[k] * (x[k] - anchor_[k]) / sides_[k]);
      if (idx[k] >= ncell_[k]) cmi_error("Position outside the box of snapshot \"%s\"!", filename_.c_str());
    }
    return values_[(idx[0] * ncell_[1] + idx[1]) * ncell_[2] + idx[2]];
  }

private:
  std::string filename_;
  double initial_neutral_fraction_;
  Vec3 anchor_, sides_;
  std::array<uint32_t, 3> ncell_;
  std::vector<DensityValues> values_;
};

/* SPH snapshot as initial condition (GadgetSnapshotDensityFunction.cpp:60-372): gas particles of a Gadget / SWIFT
 * style HDF5 snapshot (/PartType0/{Coordinates, Masses, SmoothingLength, Density, [Temperature], [NeutralFractionH]},
 * /Units, /RuntimePars:PeriodicBoundariesOn, /Header:BoxSize; fallback units from the parameter file), read with
 * host/HDF5Reader.hpp.  A cell gets the cubic-spline kernel sums at its midpoint (:315-359):
 *   density = sum_i m_i W(r_i / h_i, h_i) / 1.6737236e-27,  T = sum_i m_i W T_i / rho_i,  x_H = sum_i m_i W x_i / density
 * over the particles whose kernel contains the midpoint.  The reference finds those with an octree, one cell at a
 * time; here particles are binned on a uniform grid of the largest smoothing length, a query visits the 27 bins
 * around it (same particles, other order of the sum: rounding-level differences, tests/test_hdf5_writer.py). */
class GadgetSnapshotDensityFunction : public DensityFunction {
public:
  GadgetSnapshotDensityFunction(const std::string &name, bool fallback_periodic, double fallback_unit_length_in_SI,
                                double fallback_unit_mass_in_SI, double fallback_unit_temperature_in_SI,
                                bool use_neutral_fraction, double fallback_temperature, bool comoving_integration,
                                double hubble_parameter, Log *log = nullptr) {
    hdf5::HDF5Input file(name);
    periodic_ = fallback_periodic;
    if (file.exists("/RuntimePars")) {
      periodic_ = file.read_double_attribute("/RuntimePars", "PeriodicBoundariesOn")[0] != 0.;
    } else if (log) {
      log->write_warning("No RuntimePars found!");
    }
    Vec3 sides = {0., 0., 0.};
    if (periodic_) {
      const std::vector<double> boxsize = file.read_double_attribute("/Header", "BoxSize");
      /* a scalar BoxSize stands for a cube (HDF5Tools::read_attribute< CoordinateVector<> > needs 3 values) */
      if (boxsize.size() != 3) cmi_error("Snapshot \"%s\": /Header:BoxSize must hold 3 values!", name.c_str());
      sides = {boxsize[0], boxsize[1], boxsize[2]};
    }
    double unit_length_in_SI = fallback_unit_length_in_SI, unit_mass_in_SI = fallback_unit_mass_in_SI,
           unit_temperature_in_SI = fallback_unit_temperature_in_SI;
    if (file.exists("/Units")) {
      const double unit_length_in_cgs = file.read_double_attribute("/Units", "Unit length in cgs (U_L)")[0];
      const double unit_mass_in_cgs = file.read_double_attribute("/Units", "Unit mass in cgs (U_M)")[0];
      unit_temperature_in_SI = file.read_double_attribute("/Units", "Unit temperature in cgs (U_T)")[0];
      unit_length_in_SI = UnitConverter::to_SI(QUANTITY_LENGTH, unit_length_in_cgs, "cm");
      unit_mass_in_SI = UnitConverter::to_SI(QUANTITY_MASS, unit_mass_in_cgs, "g");
    } else {
      if (log) log->write_warning("No Units group found! Using fallback units.");
      if (unit_length_in_SI == 0.) unit_length_in_SI = 1.;
      if (unit_mass_in_SI == 0.) unit_mass_in_SI = 1.;
      if (unit_temperature_in_SI == 0.) unit_temperature_in_SI = 1.;
    }
    if (comoving_integration) {
      unit_length_in_SI /= hubble_parameter;
      unit_mass_in_SI /= hubble_parameter;
    }
    const double unit_length_in_SI_squared = unit_length_in_SI * unit_length_in_SI;
    const double unit_density_in_SI = unit_mass_in_SI / unit_length_in_SI / unit_length_in_SI_squared;
    std::vector<uint64_t> dims;
    positions_ = file.read_dataset("/PartType0/Coordinates", &dims);
    if (dims.size() != 2 || dims[1] != 3) cmi_error("Snapshot \"%s\": bad /PartType0/Coordinates!", name.c_str());
    const size_t n = dims[0];
    masses_ = file.read_dataset("/PartType0/Masses");
    smoothing_lengths_ = file.read_dataset("/PartType0/SmoothingLength");
    densities_ = file.read_dataset("/PartType0/Density");
    if (file.exists("/PartType0/Temperature")) {
      temperatures_ = file.read_dataset("/PartType0/Temperature");
    } else {
      if (fallback_temperature == 0.) fallback_temperature = 8000.;
      temperatures_.assign(n, fallback_temperature);
    }
    if (use_neutral_fraction && file.exists("/PartType0/NeutralFractionH"))
      neutral_fractions_ = file.read_dataset("/PartType0/NeutralFractionH");
    if (masses_.size() != n || smoothing_lengths_.size() != n || densities_.size() != n || temperatures_.size() != n ||
        (!neutral_fractions_.empty() && neutral_fractions_.size() != n))
      cmi_error("Snapshot \"%s\": the gas datasets have different lengths!", name.c_str());
    for (size_t i = 0; i < n; ++i) {
      for (int k = 0; k < 3; ++k) positions_[3 * i + k] *= unit_length_in_SI;
      masses_[i] *= unit_mass_in_SI;
      smoothing_lengths_[i] *= unit_length_in_SI;
      densities_[i] *= unit_density_in_SI;
      temperatures_[i] *= unit_temperature_in_SI;
    }
    for (int k = 0; k < 3; ++k) sides_[k] = sides[k] * unit_length_in_SI;
    build_bins();
  }
  explicit GadgetSnapshotDensityFunction(ParameterFile &params, Log *log = nullptr)
      : GadgetSnapshotDensityFunction(
            params.get_filename("DensityFunction:filename"),
            params.get_value<bool>("DensityFunction:fallback periodic flag", false),
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:fallback unit length", "0. m"),
            params.get_physical_value<QUANTITY_MASS>("DensityFunction:fallback unit mass", "0. kg"),
            params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:fallback unit temperature", "0. K"),
            params.get_value<bool>("DensityFunction:use neutral fraction", false),
            params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:fallback initial temperature", "0. K"),
            params.get_value<bool>("DensityFunction:comoving integration flag", false),
            params.get_value<double>("DensityFunction:hubble parameter", 0.7), log) {}

  /* CubicSplineKernel::kernel_evaluate (CubicSplineKernel.hpp:44-59) */
  static double kernel_evaluate(double u, double h) {
    const double KC1 = 2.546479089470, KC2 = 15.278874536822, KC5 = 5.092958178941;
    if (u < 1.) {
      if (u < 0.5) return (KC1 + KC2 * (u - 1.) * u * u) / (h * h * h);
      return KC5 * (1. - u) * (1. - u) * (1. - u) / (h * h * h);
    }
    return 0.;
  }
  DensityValues operator()(const Vec3 &x) override {
    double density = 0., temperature = 0., neutral_fraction = neutral_fractions_.empty() ? -1. : 0.;
    /* per axis: the bins that can hold a particle whose kernel reaches x (its own bin and the two next to it) */
    int list[3][3], nlist[3];
    for (int k = 0; k < 3; ++k) {
      const int bq = (int)std::floor((x[k] - bin_anchor_[k]) / bin_side_[k]);
      nlist[k] = 0;
      if (periodic_) {
        for (int d = -1; d <= 1; ++d) {
          const int b = wrap(bq + d, k);
          bool seen = false;
          for (int q = 0; q < nlist[k]; ++q) seen = seen || list[k][q] == b;
          if (!seen) list[k][nlist[k]++] = b;
        }
      } else if (bq >= -1 && bq <= nbin_[k] + 1) { /* the last bin also holds the particles up to the upper edge */
        const int cq = std::min(std::max(bq, 0), nbin_[k] - 1);
        for (int b = std::max(cq - 1, 0); b <= std::min(cq + 1, nbin_[k] - 1); ++b) list[k][nlist[k]++] = b;
      }
    }
    for (int a = 0; a < nlist[0]; ++a)
      for (int b = 0; b < nlist[1]; ++b)
        for (int c3 = 0; c3 < nlist[2]; ++c3) {
          const size_t bin = ((size_t)list[0][a] * nbin_[1] + list[1][b]) * nbin_[2] + list[2][c3];
          for (size_t p = bin_start_[bin]; p < bin_start_[bin + 1]; ++p) {
            const size_t i = bin_particles_[p];
            double c[3];
            for (int k = 0; k < 3; ++k) {
              c[k] = x[k] - positions_[3 * i + k];
              if (periodic_) { /* Box::periodic_distance (Box.hpp:114-127) */
                if (2 * c[k] < -sides_[k]) c[k] += sides_[k];
                if (2 * c[k] >= sides_[k]) c[k] -= sides_[k];
              }
            }
            const double r = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
            const double h = smoothing_lengths_[i];
            const double u = r / h;
            if (!(u < 1.)) continue;
            const double splineval = masses_[i] * kernel_evaluate(u, h);
            density += splineval;
            temperature += splineval * temperatures_[i] / densities_[i];
            if (neutral_fraction >= 0.) neutral_fraction += splineval * neutral_fractions_[i];
          }
        }
    DensityValues v;
    v.number_density = density / 1.6737236e-27;
    v.temperature = temperature;
    v.ionic_fraction[0] = (neutral_fraction >= 0.) ? neutral_fraction / density : 1.e-6;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }
  /* the whole grid at once: particles are scattered over the cells their kernels reach (defined in DensityGrid.hpp).
   * Same sums as operator() cell by cell, in particle order; the cost is the number of non-zero contributions,
   * whatever the spread of the smoothing lengths */
  bool set_densities(CartesianCells &grid) override;
  /* GadgetSnapshotDensityFunction::get_total_hydrogen_number (:366-372) */
  double get_total_hydrogen_number() const {
    double mtot = 0.;
    for (double m : masses_) mtot += m;
    return mtot / 1.6737236e-27;
  }
  size_t get_number_of_particles() const { return masses_.size(); }

private:
  int wrap(int b, int k) const {
    if (!periodic_) return b;
    const int n = nbin_[k];
    return ((b % n) + n) % n;
  }
  /* bins of side >= the largest smoothing length: the kernel of a particle reaches at most the neighbouring bins.
   * Periodic boxes are tiled exactly (per-axis bin side = box side / number of bins). */
  void build_bins() {
    const size_t n = masses_.size();
    double hmax = 0.;
    Vec3 lo = {DBL_MAX, DBL_MAX, DBL_MAX}, hi = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (size_t i = 0; i < n; ++i) {
      hmax = std::max(hmax, smoothing_lengths_[i]);
      for (int k = 0; k < 3; ++k) {
        lo[k] = std::min(lo[k], positions_[3 * i + k]);
        hi[k] = std::max(hi[k], positions_[3 * i + k]);
      }
    }
    if (n == 0 || !(hmax > 0.)) cmi_error("The snapshot holds no gas particles with a smoothing length!");
    if (periodic_) {
      for (int k = 0; k < 3; ++k) {
        lo[k] = 0.;
        hi[k] = sides_[k];
      }
    }
    /* at most ~8 bins per particle: memory stays O(n) when the largest kernel is tiny against the box */
    double side = hmax;
    const double volume = (hi[0] - lo[0]) * (hi[1] - lo[1]) * (hi[2] - lo[2]);
    if (volume > 0.) side = std::max(side, std::cbrt(volume / (8. * (double)n)));
    for (int k = 0; k < 3; ++k) {
      nbin_[k] = std::max(1, (int)std::floor((hi[k] - lo[k]) / side));
      bin_side_[k] = periodic_ ? sides_[k] / nbin_[k] : side;
    }
    bin_anchor_ = lo;
    const size_t nb = (size_t)nbin_[0] * nbin_[1] * nbin_[2];
    std::vector<size_t> count(nb + 1, 0), which(n);
    for (size_t i = 0; i < n; ++i) {
      int b[3];
      for (int k = 0; k < 3; ++k) {
        b[k] = (int)std::floor((positions_[3 * i + k] - bin_anchor_[k]) / bin_side_[k]);
        b[k] = periodic_ ? wrap(b[k], k) : std::min(std::max(b[k], 0), nbin_[k] - 1);
      }
      which[i] = ((size_t)b[0] * nbin_[1] + b[1]) * nbin_[2] + b[2];
      ++count[which[i] + 1];
    }
    for (size_t b = 0; b < nb; ++b) count[b + 1] += count[b];
    bin_start_ = count;
    bin_particles_.resize(n);
    std::vector<size_t> fill(bin_start_.begin(), bin_start_.end() - 1);
    for (size_t i = 0; i < n; ++i) bin_particles_[fill[which[i]]++] = i;
  }

  bool periodic_ = false;
  Vec3 sides_ = {0., 0., 0.};
  std::vector<double> positions_, masses_, smoothing_lengths_, densities_, temperatures_, neutral_fractions_;
  Vec3 bin_side_ = {0., 0., 0.}, bin_anchor_ = {0., 0., 0.};
  int nbin_[3] = {1, 1, 1};
  std::vector<size_t> bin_start_, bin_particles_;
};

/* Lambert W function on (-1/e, 0), branches 0 and -1 (LambertW.hpp:44-110): a 5th order series as first guess
 * (mirrored around W = -1 for the lower branch), then Newton steps to a relative tolerance of 1e-10 */
inline double lambert_w(double r, int branch) {
  if (r >= 0. || r < -1. / M_E) cmi_error("Input value for Lambert W outside supported range: %g!", r);
  const double r2 = r * r, r3 = r2 * r, r4 = r2 * r2, r5 = r4 * r;
  const double w = r - r2 + 1.5 * r3 - (8. / 3.) * r4 + (125. / 24.) * r5;
  auto newton_step = [r](double w_) {
    const double expw = std::exp(w_);
    const double wexpw = w_ * expw;
    return w_ - (wexpw - r) / (expw + wexpw);
  };
  double w0 = (branch == 0) ? w : -2. - w;
  double w1 = newton_step(w0);
  while (std::abs(w0 - w1) > std::abs(w0 + w1) * 1.e-10) {
    w0 = w1;
    w1 = newton_step(w0);
  }
  return w1;
}

/* Spherical Bondi accretion onto a point mass, optionally with an ionised inner region in pressure balance
 * (BondiProfile.hpp:82-236, BondiProfileDensityFunction.hpp:52-117): density and pressure of the analytic flow,
 * temperature = m_p P / (k rho), halved where the profile is ionised.  The velocities of the profile belong to
 * the hydro and are not part of the grid here. */
class BondiProfileDensityFunction : public DensityFunction {
public:
  BondiProfileDensityFunction(double central_mass, double bondi_density, double sound_speed, double ionisation_radius,
                              double pressure_contrast, const Vec3 &center, double neutral_fraction)
      : bondi_radius_(0.5 * constants::newton_constant * central_mass / (sound_speed * sound_speed)),
        bondi_density_(bondi_density), sound_speed_(sound_speed), ionisation_radius_(ionisation_radius),
        pressure_contrast_(pressure_contrast), center_(center), neutral_fraction_(neutral_fraction) {
    if (ionisation_radius_ > 0. && pressure_contrast_ > 0.) {
      const double rBI = bondi_radius_ / ionisation_radius_;
      const double rBI2 = rBI * rBI;
      const double lambertarg = -rBI2 * rBI2 * std::exp(3. - 4. * rBI);
      double v_RI = std::sqrt(-lambert_w(lambertarg, -1));
      const double rho_RI = rBI2 * bondi_density_ / v_RI;
      v_RI *= -sound_speed_;
      const double vRI2_Pccs2 = v_RI * v_RI / (pressure_contrast_ * sound_speed_ * sound_speed_);
      const double Pc_inv = 1. / pressure_contrast_;
      const double sum = vRI2_Pccs2 + Pc_inv;
      const double Gamma = 0.5 * (sum - std::sqrt(sum * sum - 4. * vRI2_Pccs2));
      rho_I_ = Gamma * rho_RI;
      v_I_ = v_RI / Gamma;
    }
  }
  explicit BondiProfileDensityFunction(ParameterFile &params)
      : BondiProfileDensityFunction(
            params.get_physical_value<QUANTITY_MASS>("DensityFunction:central mass", "18. Msol"),
            params.get_physical_value<QUANTITY_DENSITY>("DensityFunction:Bondi density", "1.e-19 g cm^-3"),
            params.get_physical_value<QUANTITY_VELOCITY>("DensityFunction:sound speed", "2.031 km s^-1"),
            params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:ionisation radius", "0. m"),
            params.get_value<double>("DensityFunction:pressure contrast", 32.),
            params.get_physical_vector<QUANTITY_LENGTH>("DensityFunction:center", "[0. m, 0. m, 0. m]"),
            params.get_value<double>("DensityFunction:neutral fraction", 1.)) {
    params.get_physical_value<QUANTITY_LENGTH>("DensityFunction:vprof radius", "0. m");
    params.get_physical_value<QUANTITY_VELOCITY>("DensityFunction:vprof velocity", "0. m s^-1");
  }
  DensityValues operator()(const Vec3 &x) override {
    const Vec3 relpos = {x[0] - center_[0], x[1] - center_[1], x[2] - center_[2]};
    const double radius = std::sqrt(relpos[0] * relpos[0] + relpos[1] * relpos[1] + relpos[2] * relpos[2]);
    const double inverse_radius = 1. / radius;
    double density, pressure, profile_neutral_fraction;
    const double rB = bondi_radius_ * inverse_radius;
    if (rB < 184.5) { /* the solution diverges for very small radii (BondiProfile.hpp:171-175) */
      const double rB2 = rB * rB;
      const double lambertarg = -rB2 * rB2 * std::exp(3. - 4. * rB);
      double v_cs;
      if (radius > bondi_radius_) {
        v_cs = std::sqrt(-lambert_w(lambertarg, 0));
      } else if (radius < ionisation_radius_) {
        const double RIr = ionisation_radius_ * inverse_radius;
        const double RIr2 = RIr * RIr;
        const double vI2_Pccs2 = v_I_ * v_I_ / (pressure_contrast_ * sound_speed_ * sound_speed_);
        const double lambertarg2 =
            -RIr2 * RIr2 * vI2_Pccs2 *
            std::exp(4. * bondi_radius_ / pressure_contrast_ * (1. / ionisation_radius_ - inverse_radius) - vI2_Pccs2);
        v_cs = std::sqrt(-pressure_contrast_ * lambert_w(lambertarg2, -1));
      } else {
        v_cs = std::sqrt(-lambert_w(lambertarg, -1));
      }
      const double vB = -v_cs * sound_speed_;
      if (radius < ionisation_radius_) {
        density = rho_I_ * ionisation_radius_ * ionisation_radius_ * v_I_ / (radius * radius * vB);
        pressure = sound_speed_ * sound_speed_ * pressure_contrast_ * density;
        profile_neutral_fraction = 0.;
      } else {
        density = rB2 * bondi_density_ / v_cs;
        pressure = sound_speed_ * sound_speed_ * density;
        profile_neutral_fraction = 1.;
      }
    } else {
      density = bondi_density_;
      pressure = sound_speed_ * sound_speed_ * density;
      profile_neutral_fraction = 1.;
    }
    DensityValues v;
    v.number_density = density / constants::proton_mass;
    double temperature = constants::proton_mass * pressure / (constants::boltzmann * density);
    if (profile_neutral_fraction < 0.5) temperature *= 0.5; /* ionised gas has a lower mean molecular mass */
    v.temperature = temperature;
    v.ionic_fraction[0] = neutral_fraction_;
    v.ionic_fraction[1] = 1.e-6;
    return v;
  }

private:
  double bondi_radius_, bondi_density_, sound_speed_, ionisation_radius_, pressure_contrast_;
  double rho_I_ = 0., v_I_ = 0.;
  Vec3 center_;
  double neutral_fraction_;
};

/* Block-structured AMR snapshot of FLASH (FLASHSnapshotDensityFunction.cpp:60-215, :228-241): the leaf blocks
 * (node type 1) of "bounding box" / "dens" / "temp" with the box and block counts of the runtime parameter tables,
 * CGS units; a position gets the values of the FLASH cell that contains it (the reference inserts every cell into
 * an AMR tree by the key of its centre and looks positions up in that tree: the same cell).  `temperature` > 0
 * overrides the snapshot's temperatures; the cosmic-ray heating variant needs the AMR neighbour search and is refused. */
class FLASHSnapshotDensityFunction : public DensityFunction {
public:
  FLASHSnapshotDensityFunction(std::string filename, double temperature, bool read_cosmic_ray_heating)
      : filename_(std::move(filename)), temperature_(temperature) {
    if (read_cosmic_ray_heating)
      cmi_error("DensityFunction:read cosmic ray heating is not provided by the B200 backend!");
  }
  explicit FLASHSnapshotDensityFunction(ParameterFile &params)
      : FLASHSnapshotDensityFunction(params.get_filename("DensityFunction:filename"),
                                     params.get_physical_value<QUANTITY_TEMPERATURE>("DensityFunction:temperature", "-1. K"),
                                     params.get_value<bool>("DensityFunction:read cosmic ray heating", false)) {}
  void initialize() override {
    hdf5::HDF5Input file(filename_);
    const double unit_length_in_SI = UnitConverter::to_SI(QUANTITY_LENGTH, 1., "cm");
    const double unit_density_in_SI = UnitConverter::to_SI(QUANTITY_DENSITY, 1., "g cm^-3");
    const std::map<std::string, double> real_pars = file.read_dictionary("real runtime parameters");
    auto par = [&](const char *name) {
      const auto it = real_pars.find(name);
      if (it == real_pars.end()) cmi_error("Snapshot \"%s\": no runtime parameter \"%s\"!", filename_.c_str(), name);
      return it->second;
    };
    const char *lo_names[3] = {"xmin", "ymin", "zmin"}, *hi_names[3] = {"xmax", "ymax", "zmax"};
    for (int d = 0; d < 3; ++d) {
      anchor_[d] = par(lo_names[d]) * unit_length_in_SI;
      top_[d] = par(hi_names[d]) * unit_length_in_SI;
    }
    std::vector<uint64_t> edims, ddims;
    const std::vector<double> extents = file.read_dataset("bounding box", &edims);
    densities_ = file.read_dataset("dens", &ddims);
    if (temperature_ <= 0.) temperatures_ = file.read_dataset("temp");
    const std::vector<double> nodetypes = file.read_dataset("node type");
    if (edims.size() != 3 || edims[1] != 3 || edims[2] != 2 || ddims.size() != 4 || ddims[0] != edims[0] ||
        nodetypes.size() != edims[0] || (temperature_ <= 0. && temperatures_.size() != densities_.size()))
      cmi_error("Snapshot \"%s\" does not have the layout of a FLASH file!", filename_.c_str());
    for (int d = 0; d < 3; ++d) ncell_[d] = ddims[3 - d]; /* "dens" is [block][z][y][x] */
    for (size_t i = 0; i < edims[0]; ++i) {
      if (nodetypes[i] != 1.) continue;
      Block b;
      b.index = i;
      for (int d = 0; d < 3; ++d) {
        b.anchor[d] = extents[(i * 3 + d) * 2] * unit_length_in_SI;
        b.sides[d] = extents[(i * 3 + d) * 2 + 1] * unit_length_in_SI - b.anchor[d];
      }
      blocks_.push_back(b);
    }
    if (blocks_.empty()) cmi_error("Snapshot \"%s\" holds no leaf blocks!", filename_.c_str());
    for (double &rho : densities_) rho *= unit_density_in_SI;
  }
  DensityValues operator()(const Vec3 &x) override {
    for (const Block &b : blocks_) {
      size_t c[3];
      bool inside = true;
      for (int d = 0; d < 3 && inside; ++d) {
        const double f = (x[d] - b.anchor[d]) / b.sides[d];
        inside = f >= 0. && f < 1.;
        if (inside) c[d] = std::min((size_t)(f * ncell_[d]), (size_t)ncell_[d] - 1);
      }
      if (!inside) continue;
      const size_t at = ((b.index * ncell_[2] + c[2]) * ncell_[1] + c[1]) * ncell_[0] + c[0];
      DensityValues v;
      v.number_density = densities_[at] / 1.6737236e-27;
      v.temperature = (temperature_ <= 0.) ? temperatures_[at] : temperature_;
      v.ionic_fraction[0] = 1.e-6;
      v.ionic_fraction[1] = 1.e-6;
      return v;
    }
    cmi_error("Position [%g m, %g m, %g m] lies outside the blocks of snapshot \"%s\"!", x[0], x[1], x[2], filename_.c_str());
  }
  size_t get_number_of_leaf_blocks() const { return blocks_.size(); }
  /* the whole grid at once: every leaf block hands its cells to the grid cells whose midpoints it contains (defined
   * in DensityGrid.hpp); same cells as operator(), without a search over the blocks per cell */
  bool set_densities(CartesianCells &grid) override;

private:
  struct Block {
    size_t index;
    double anchor[3], sides[3];
  };
  std::string filename_;
  double temperature_;
  Vec3 anchor_ = {0., 0., 0.}, top_ = {0., 0., 0.};
  uint64_t ncell_[3] = {1, 1, 1};
  std::vector<Block> blocks_;
  std::vector<double> densities_, temperatures_;
};

struct DensityFunctionFactory {
  static DensityFunction *generate(ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("DensityFunction:type", "Homogeneous");
    if (log) log->write_info("Requested DensityFunction type: ", type);
    if (type == "Homogeneous") return new HomogeneousDensityFunction(params);
    if (type == "BlockSyntax") return new BlockSyntaxDensityFunction(params);
    if (type == "AsciiFile") return new AsciiFileDensityFunction(params);
    if (type == "Interpolated") return new InterpolatedDensityFunction(params);
    if (type == "BondiProfile") return new BondiProfileDensityFunction(params);
    if (type == "CoredDMProfile") return new CoredDMProfileDensityFunction(params);
    if (type == "DiscIC") return new DiscICDensityFunction(params);
    if (type == "DiscPatch") return new DiscPatchDensityFunction(params);
    if (type == "SpiralGalaxy") return new SpiralGalaxyDensityFunction(params);
    if (type == "CMacIonizeSnapshot") return new CMacIonizeSnapshotDensityFunction(params);
    if (type == "GadgetSnapshot") return new GadgetSnapshotDensityFunction(params, log);
    if (type == "FLASHSnapshot") return new FLASHSnapshotDensityFunction(params);
    cmi_error("Unknown DensityFunction type: \"%s\" (the B200 backend provides Homogeneous, BlockSyntax, AsciiFile, "
              "Interpolated, BondiProfile, CoredDMProfile, DiscIC, DiscPatch, SpiralGalaxy, CMacIonizeSnapshot, GadgetSnapshot and FLASHSnapshot)!",
              type.c_str());
  }
};

} // namespace cmi
