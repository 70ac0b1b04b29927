/*
 * IonizationSimulation.hpp — C++ host layer of the B200 backend: the reference's
 * plugin classes for the photoionization path, re-implemented as thin owners of
 * parameters that configure one `cmib_context` (include/cmib.h) per GPU.
 *
 * Same class names, parameter keys, defaults and error messages as the reference so
 * that an existing parameter file and an existing caller keep working:
 *
 *   this file                       reference (under /root/reference/src)
 *   ------------------------------  -------------------------------------------------
 *   SimulationBox                   SimulationBox.hpp:63-72
 *   DensityFunction (+Factory)      DensityFunctionFactory.hpp; HomogeneousDensityFunction.hpp:83-108;
 *                                   BlockSyntaxDensityFunction.hpp:75-199, BlockSyntaxBlock.hpp:91-106
 *   PhotonSourceDistribution        PhotonSourceDistributionFactory.hpp:99; SingleStarPhotonSourceDistribution.hpp:78-85;
 *                                   AsciiFileTablePhotonSourceDistribution.cpp:40-118
 *   PhotonSourceSpectrum            PhotonSourceSpectrumFactory.hpp:84-152; Monochromatic...hpp:78-86; Planck...cpp:128-139
 *   CrossSections                   CrossSectionsFactory.hpp:60-80; FixedValueCrossSections.hpp:112-141
 *   RecombinationRates              RecombinationRatesFactory.hpp:59-72; FixedValueRecombinationRates.hpp:118-147
 *   AbundanceModel                  AbundanceModelFactory.hpp:54-89; FixedValueAbundanceModel.hpp:54-60
 *   DiffuseReemissionHandler        DiffuseReemissionHandlerFactory.hpp:59-107; FixedValueDiffuseReemissionHandler.hpp:66-72
 *   TemperatureCalculator params    TemperatureCalculator.cpp:133-160
 *   CartesianDensityGrid            CartesianDensityGrid.cpp:44-134, .hpp:85-144; DensityGrid.hpp:235-259,775-790
 *   AsciiFileDensityGridWriter      AsciiFileDensityGridWriter.cpp:58-95
 *   IonizationSimulation            IonizationSimulation.cpp:101-231 (ctor), :239-326 (initialize), :334-679 (run)
 *
 * What is NOT here: every compute step.  Emission, the voxel walk, accumulation,
 * re-emission and the per-cell ionization/temperature solve run on the GPU behind
 * the C ABI; this layer only builds inputs, orders the calls of one iteration and
 * moves results.  There is no CPU compute path.
 */
#pragma once
#include <array>
#include <cfloat>
#include <chrono>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <dlfcn.h>
#include <sys/utsname.h>
#include <nccl.h> /* types only: the library is bound at run time, see NcclApi */

#include "../../include/cmib.h"
#include "Error.hpp"
#include "HDF5Reader.hpp"
#include "HDF5Writer.hpp"
#include "ParameterFile.hpp"
#include "RandomGenerator.hpp"
#include "../csrc/spectrum_tables.hpp" /* host-side table builders + the samplers the device uses (plain C++) */

namespace cmi {

using Vec3 = std::array<double, 3>;

/* ---- logging: same levels as the reference's Log (Log.hpp:41-46), terminal only ---- */
class Log {
public:
  enum Level { INFO = 0, STATUS, WARNING, ERROR_ };
  explicit Log(Level level = STATUS, std::ostream &out = std::cerr) : level_(level), out_(out) {}
  template <class... A> void write_info(const A &...a) { write(INFO, a...); }
  template <class... A> void write_status(const A &...a) { write(STATUS, a...); }
  template <class... A> void write_warning(const A &...a) { write(WARNING, a...); }

private:
  Level level_;
  std::ostream &out_;
  template <class... A> void write(Level l, const A &...a) {
    if (l < level_) return;
    std::ostringstream s;
    (void)std::initializer_list<int>{(s << a, 0)...};
    out_ << s.str() << "\n";
  }
};

#define CMIB_CALL(expr)                                                                         \
  do {                                                                                          \
    if ((expr) != 0) cmi_error("%s failed: %s", #expr, cmib_last_error());                      \
  } while (0)

/* ---- ion / element names of the parameter files (ElementNames.hpp:107-160, 52-88) ---- */
inline const char *ion_name(int ion) {
  static const char *names[CMIB_NUM_IONS] = {"H_n", "He_n", "C_p1", "C_p2", "N_n", "N_p1", "N_p2",
                                             "O_n", "O_p1", "Ne_n", "Ne_p1", "S_p1", "S_p2", "S_p3"};
  return names[ion];
}
/* get_ion_name (ElementNames.hpp:210-240): the names snapshot fields carry (NeutralFractionH, NeutralFractionC+, ...) */
inline const char *ion_symbol(int ion) {
  static const char *names[CMIB_NUM_IONS] = {"H", "He", "C+", "C++", "N", "N+", "N++", "O", "O+", "Ne", "Ne+", "S+", "S++", "S+++"};
  return names[ion];
}
inline const char *element_name(int el) {
  static const char *names[CMIB_NUM_ELEMENTS] = {"He", "C", "N", "O", "Ne", "S"};
  return names[el];
}

struct SimulationBox {
  Vec3 anchor, sides;
  std::array<bool, 3> periodicity;
  explicit SimulationBox(ParameterFile &params)
      : anchor(params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:anchor", "[-5. pc, -5. pc, -5. pc]")),
        sides(params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:sides", "[10. pc, 10. pc, 10. pc]")),
        periodicity(params.get_value<std::array<bool, 3>>("SimulationBox:periodicity", {false, false, false})) {}
};

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

struct DensityFunctionFactory {
  static DensityFunction *generate(ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("DensityFunction:type", "Homogeneous");
    if (log) log->write_info("Requested DensityFunction type: ", type);
    if (type == "Homogeneous") return new HomogeneousDensityFunction(params);
    if (type == "BlockSyntax") return new BlockSyntaxDensityFunction(params);
    if (type == "AsciiFile") return new AsciiFileDensityFunction(params);
    if (type == "Interpolated") return new InterpolatedDensityFunction(params);
    if (type == "CoredDMProfile") return new CoredDMProfileDensityFunction(params);
    if (type == "DiscIC") return new DiscICDensityFunction(params);
    if (type == "DiscPatch") return new DiscPatchDensityFunction(params);
    if (type == "SpiralGalaxy") return new SpiralGalaxyDensityFunction(params);
    if (type == "CMacIonizeSnapshot") return new CMacIonizeSnapshotDensityFunction(params);
    if (type == "GadgetSnapshot") return new GadgetSnapshotDensityFunction(params, log);
    cmi_error("Unknown DensityFunction type: \"%s\" (the B200 backend provides Homogeneous, BlockSyntax, AsciiFile, "
              "Interpolated, CoredDMProfile, DiscIC, DiscPatch, SpiralGalaxy, CMacIonizeSnapshot and GadgetSnapshot)!",
              type.c_str());
  }
};

/* ---- PhotonSourceDistribution ---- */
class PhotonSourceDistribution {
public:
  virtual ~PhotonSourceDistribution() {}
  virtual size_t get_number_of_sources() const = 0;
  virtual Vec3 get_position(size_t index) = 0;
  virtual double get_weight(size_t index) const = 0;
  virtual double get_total_luminosity() const = 0;
};

class SingleStarPhotonSourceDistribution : public PhotonSourceDistribution {
public:
  SingleStarPhotonSourceDistribution(const Vec3 &position, double luminosity)
      : position_(position), luminosity_(luminosity) {}
  explicit SingleStarPhotonSourceDistribution(ParameterFile &params)
      : SingleStarPhotonSourceDistribution(
            params.get_physical_vector<QUANTITY_LENGTH>("PhotonSourceDistribution:position", "[0. pc, 0. pc, 0. pc]"),
            params.get_physical_value<QUANTITY_FREQUENCY>("PhotonSourceDistribution:luminosity", "4.26e49 s^-1")) {}
  size_t get_number_of_sources() const override { return 1; }
  Vec3 get_position(size_t) override { return position_; }
  double get_weight(size_t) const override { return 1.; }
  double get_total_luminosity() const override { return luminosity_; }

private:
  Vec3 position_;
  double luminosity_;
};

/* number of sources, total luminosity, then "x y z weight" rows (SI); '#' comments */
class AsciiFileTablePhotonSourceDistribution : public PhotonSourceDistribution {
public:
  explicit AsciiFileTablePhotonSourceDistribution(const std::string &filename) {
    std::ifstream file(filename);
    if (!file.is_open()) cmi_error("Could not open file \"%s\"!", filename.c_str());
    std::string line;
    size_t n = 0, got = 0;
    int stage = 0;
    while (std::getline(file, line)) {
      if (line.empty() || line[0] == '#') continue;
      std::stringstream ls(line);
      if (stage == 0) {
        ls >> n;
        positions_.resize(n);
        weights_.resize(n);
        stage = 1;
      } else if (stage == 1) {
        ls >> luminosity_;
        stage = 2;
      } else {
        if (got == n) break;
        ls >> positions_[got][0] >> positions_[got][1] >> positions_[got][2] >> weights_[got];
        ++got;
      }
    }
    if (got < n) cmi_error("The file %s has fewer sources (%zu) than needed (%zu).\n", filename.c_str(), got, n);
  }
  explicit AsciiFileTablePhotonSourceDistribution(ParameterFile &params)
      : AsciiFileTablePhotonSourceDistribution(
            params.get_value<std::string>("PhotonSourceDistribution:filename", "sinks.txt")) {}
  size_t get_number_of_sources() const override { return positions_.size(); }
  Vec3 get_position(size_t i) override { return positions_[i]; }
  double get_weight(size_t i) const override { return weights_[i]; }
  double get_total_luminosity() const override { return luminosity_; }

private:
  std::vector<Vec3> positions_;
  std::vector<double> weights_;
  double luminosity_ = 0.;
};

/* AsciiFilePhotonSourceDistribution (src/AsciiFilePhotonSourceDistribution.hpp:50-98): a YAML file with
 * "number of sources" and source[i]:position / source[i]:luminosity */
class AsciiFilePhotonSourceDistribution : public PhotonSourceDistribution {
public:
  explicit AsciiFilePhotonSourceDistribution(const std::string &filename) {
    std::ifstream file(filename);
    if (!file) cmi_error("Error while opening file \"%s\"!", filename.c_str());
    YAMLDictionary blocks(file);
    const uint32_t n = blocks.get_value<uint32_t>("number of sources");
    positions_.resize(n);
    luminosities_.resize(n);
    for (uint32_t i = 0; i < n; ++i) {
      const std::string name = "source[" + std::to_string(i) + "]:";
      positions_[i] = blocks.get_physical_vector<QUANTITY_LENGTH>(name + "position");
      luminosities_[i] = blocks.get_physical_value<QUANTITY_FREQUENCY>(name + "luminosity");
      total_luminosity_ += luminosities_[i];
    }
    std::ofstream ofile(filename + ".used-values");
    blocks.print_contents(ofile, true);
  }
  explicit AsciiFilePhotonSourceDistribution(ParameterFile &params)
      : AsciiFilePhotonSourceDistribution(params.get_filename("PhotonSourceDistribution:filename", "sources.yml")) {}
  size_t get_number_of_sources() const override { return positions_.size(); }
  Vec3 get_position(size_t i) override { return positions_[i]; }
  double get_weight(size_t i) const override { return luminosities_[i] / total_luminosity_; }
  double get_total_luminosity() const override { return total_luminosity_; }

private:
  std::vector<Vec3> positions_;
  std::vector<double> luminosities_;
  double total_luminosity_ = 0.;
};

/* UniformRandomPhotonSourceDistribution (src/UniformRandomPhotonSourceDistribution.hpp:88-300): equal
 * sources at positions drawn uniformly in a box with the reference's generator (RandomGenerator.hpp:
 * same seed, same positions), each with a random remaining lifetime; the population is evolved in
 * steps of the update interval up to the starting time (dead sources are replaced).  The
 * time-dependent update() belongs to the radiation-hydrodynamics driver and is not provided. */
class UniformRandomPhotonSourceDistribution : public PhotonSourceDistribution {
public:
  UniformRandomPhotonSourceDistribution(double source_lifetime, double source_luminosity, uint32_t number_of_sources,
                                        const Vec3 &box_anchor, const Vec3 &box_sides, int32_t seed,
                                        double update_interval, double starting_time)
      : source_luminosity_(source_luminosity), anchor_(box_anchor), sides_(box_sides), random_generator_(seed) {
    for (uint32_t i = 0; i < number_of_sources; ++i) {
      lifetimes_.push_back(random_generator_.get_uniform_random_double() * source_lifetime);
      positions_.push_back(generate_source_position());
    }
    uint32_t number_of_updates = 1;
    while (number_of_updates * update_interval <= starting_time) {
      size_t i = 0;
      while (i < lifetimes_.size()) {
        lifetimes_[i] -= update_interval;
        if (lifetimes_[i] <= 0.) {
          positions_.erase(positions_.begin() + i);
          lifetimes_.erase(lifetimes_.begin() + i);
        } else {
          ++i;
        }
      }
      for (size_t k = positions_.size(); k < number_of_sources; ++k) {
        const double offset = random_generator_.get_uniform_random_double() * update_interval;
        lifetimes_.push_back(source_lifetime - offset);
        positions_.push_back(generate_source_position());
      }
      ++number_of_updates;
    }
  }
  explicit UniformRandomPhotonSourceDistribution(ParameterFile &params)
      : UniformRandomPhotonSourceDistribution(
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:source lifetime", "1. Myr"),
            params.get_physical_value<QUANTITY_FREQUENCY>("PhotonSourceDistribution:source luminosity", "1.e48 s^-1"),
            params.get_value<uint32_t>("PhotonSourceDistribution:number of sources", 1),
            params.get_physical_vector<QUANTITY_LENGTH>("PhotonSourceDistribution:box anchor", "[-5. pc, -5. pc, -5. pc]"),
            params.get_physical_vector<QUANTITY_LENGTH>("PhotonSourceDistribution:box sides", "[10. pc, 10. pc, 10. pc]"),
            params.get_value<int32_t>("PhotonSourceDistribution:random seed", 42),
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:update interval", "0.1 Myr"),
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:starting time", "0. Myr")) {
    if (params.get_value<bool>("PhotonSourceDistribution:output sources", false))
      cmi_error("PhotonSourceDistribution:output sources is not provided by the B200 backend!");
  }
  size_t get_number_of_sources() const override { return positions_.size(); }
  Vec3 get_position(size_t i) override { return positions_[i]; }
  double get_weight(size_t) const override { return 1. / get_number_of_sources(); }
  double get_total_luminosity() const override { return source_luminosity_ * get_number_of_sources(); }

private:
  Vec3 generate_source_position() {
    Vec3 p;
    for (int d = 0; d < 3; ++d) p[d] = anchor_[d] + random_generator_.get_uniform_random_double() * sides_[d];
    return p;
  }
  double source_luminosity_;
  Vec3 anchor_, sides_;
  RandomGenerator random_generator_;
  std::vector<Vec3> positions_;
  std::vector<double> lifetimes_;
};

/* Sources that are born at random and die after a fixed lifetime, evolved in steps of the update
 * interval up to the starting time: every step each of `average_number` slots gives birth with
 * probability update interval / lifetime (DiscPatchPhotonSourceDistribution.hpp:131-204,
 * DwarfGalaxyPhotonSourceDistribution.hpp:126-197: the two differ in where a source is put). */
class StochasticPhotonSourcePopulation : public PhotonSourceDistribution {
public:
  size_t get_number_of_sources() const override { return positions_.size(); }
  Vec3 get_position(size_t i) override { return positions_[i]; }
  double get_weight(size_t) const override { return 1. / get_number_of_sources(); }
  double get_total_luminosity() const override { return source_luminosity_ * get_number_of_sources(); }

protected:
  StochasticPhotonSourcePopulation(double source_luminosity, int32_t seed)
      : source_luminosity_(source_luminosity), random_generator_(seed) {}
  virtual Vec3 generate_source_position() = 0;
  /* called by the concrete class once its position parameters are in place */
  void populate(double source_lifetime, uint32_t average_number, double update_interval, double starting_time) {
    const double source_probability = update_interval / source_lifetime;
    for (uint32_t i = 0; i < average_number; ++i) {
      lifetimes_.push_back(random_generator_.get_uniform_random_double() * source_lifetime);
      positions_.push_back(generate_source_position());
    }
    uint32_t number_of_updates = 1;
    while (number_of_updates * update_interval <= starting_time) {
      size_t i = 0;
      while (i < lifetimes_.size()) {
        lifetimes_[i] -= update_interval;
        if (lifetimes_[i] <= 0.) {
          positions_.erase(positions_.begin() + i);
          lifetimes_.erase(lifetimes_.begin() + i);
        } else {
          ++i;
        }
      }
      for (uint32_t k = 0; k < average_number; ++k) {
        if (random_generator_.get_uniform_random_double() <= source_probability) {
          const double offset = random_generator_.get_uniform_random_double() * update_interval;
          lifetimes_.push_back(source_lifetime - offset);
          positions_.push_back(generate_source_position());
        }
      }
      ++number_of_updates;
    }
  }
  static void no_source_output(ParameterFile &params) {
    if (params.get_value<bool>("PhotonSourceDistribution:output sources", false))
      cmi_error("PhotonSourceDistribution:output sources is not provided by the B200 backend!");
  }
  /* one Box-Muller deviate: scale * sqrt(-2 ln u1) * cos(2 pi u2) */
  double gaussian(double scale) {
    const double rho = scale * std::sqrt(-2. * std::log(random_generator_.get_uniform_random_double()));
    return rho * std::cos(2. * M_PI * random_generator_.get_uniform_random_double());
  }
  double source_luminosity_;
  RandomGenerator random_generator_;
  std::vector<Vec3> positions_;
  std::vector<double> lifetimes_;
};

/* uniform in x and y over a rectangle, Gaussian in z */
class DiscPatchPhotonSourceDistribution : public StochasticPhotonSourcePopulation {
public:
  DiscPatchPhotonSourceDistribution(double source_lifetime, double source_luminosity, uint32_t average_number,
                                    double anchor_x, double sides_x, double anchor_y, double sides_y, double origin_z,
                                    double scaleheight_z, int32_t seed, double update_interval, double starting_time)
      : StochasticPhotonSourcePopulation(source_luminosity, seed), anchor_x_(anchor_x), sides_x_(sides_x),
        anchor_y_(anchor_y), sides_y_(sides_y), origin_z_(origin_z), scaleheight_z_(scaleheight_z) {
    populate(source_lifetime, average_number, update_interval, starting_time);
  }
  explicit DiscPatchPhotonSourceDistribution(ParameterFile &params)
      : DiscPatchPhotonSourceDistribution(
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:source lifetime", "20. Myr"),
            params.get_physical_value<QUANTITY_FREQUENCY>("PhotonSourceDistribution:source luminosity", "3.125e49 s^-1"),
            params.get_value<uint32_t>("PhotonSourceDistribution:average number of sources", 24),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:anchor x", "-1. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:sides x", "2. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:anchor y", "-1. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:sides y", "2. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:origin z", "0. pc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:scaleheight z", "63. pc"),
            params.get_value<int32_t>("PhotonSourceDistribution:random seed", 42),
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:update interval", "0.1 Myr"),
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:starting time", "0. Myr")) {
    no_source_output(params);
  }

protected:
  Vec3 generate_source_position() override {
    Vec3 p;
    p[0] = anchor_x_ + random_generator_.get_uniform_random_double() * sides_x_;
    p[1] = anchor_y_ + random_generator_.get_uniform_random_double() * sides_y_;
    p[2] = gaussian(scaleheight_z_) + origin_z_;
    return p;
  }

private:
  double anchor_x_, sides_x_, anchor_y_, sides_y_, origin_z_, scaleheight_z_;
};

/* Gaussian blob: (x, y) from one Box-Muller pair, z from a second one.  The reference reads a
 * `center` but never adds it to the positions (DwarfGalaxyPhotonSourceDistribution.hpp:98-119);
 * neither does this class. */
class DwarfGalaxyPhotonSourceDistribution : public StochasticPhotonSourcePopulation {
public:
  DwarfGalaxyPhotonSourceDistribution(double source_lifetime, double source_luminosity, uint32_t average_number,
                                      double scale_radius, int32_t seed, double update_interval, double starting_time)
      : StochasticPhotonSourcePopulation(source_luminosity, seed), scale_radius_(scale_radius) {
    populate(source_lifetime, average_number, update_interval, starting_time);
  }
  explicit DwarfGalaxyPhotonSourceDistribution(ParameterFile &params)
      : DwarfGalaxyPhotonSourceDistribution(
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:source lifetime", "20. Myr"),
            params.get_physical_value<QUANTITY_FREQUENCY>("PhotonSourceDistribution:source luminosity", "3.125e49 s^-1"),
            params.get_value<uint32_t>("PhotonSourceDistribution:average number of sources", 52),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:scale radius", "300. pc"),
            params.get_value<int32_t>("PhotonSourceDistribution:random seed", 42),
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:update interval", "0.01 Gyr"),
            params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:starting time", "0. Gyr")) {
    params.get_physical_vector<QUANTITY_LENGTH>("PhotonSourceDistribution:center", "[0. kpc, 0. kpc, 0. kpc]");
    no_source_output(params);
  }

protected:
  Vec3 generate_source_position() override {
    const double rho1 = scale_radius_ * std::sqrt(-2. * std::log(random_generator_.get_uniform_random_double()));
    const double phi1 = 2. * M_PI * random_generator_.get_uniform_random_double();
    const double rho2 = scale_radius_ * std::sqrt(-2. * std::log(random_generator_.get_uniform_random_double()));
    const double phi2 = 2. * M_PI * random_generator_.get_uniform_random_double();
    return Vec3{rho1 * std::cos(phi1), rho1 * std::sin(phi1), rho2 * std::cos(phi2)};
  }

private:
  double scale_radius_;
};

/* A fixed number of equal sources, uniform in x and y, Gaussian in z; a position is drawn when
 * it is asked for (SILCCPhotonSourceDistribution.hpp:159-187), so asking twice gives two answers. */
class SILCCPhotonSourceDistribution : public PhotonSourceDistribution {
public:
  SILCCPhotonSourceDistribution(uint32_t num_sources, double anchor_x, double sides_x, double anchor_y, double sides_y,
                                double origin_z, double scaleheight_z, double luminosity, int32_t seed)
      : num_sources_(num_sources), anchor_x_(anchor_x), sides_x_(sides_x), anchor_y_(anchor_y), sides_y_(sides_y),
        origin_z_(origin_z), scaleheight_z_(scaleheight_z), luminosity_(luminosity), random_generator_(seed) {}
  explicit SILCCPhotonSourceDistribution(ParameterFile &params)
      : SILCCPhotonSourceDistribution(
            params.get_value<uint32_t>("PhotonSourceDistribution:number of sources", 24),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:anchor x", "-1. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:sides x", "2. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:anchor y", "-1. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:sides y", "2. kpc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:origin z", "0. pc"),
            params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:scaleheight z", "63. pc"),
            params.get_physical_value<QUANTITY_FREQUENCY>("PhotonSourceDistribution:luminosity", "3.125e49 s^-1"),
            params.get_value<int32_t>("PhotonSourceDistribution:random seed", 42)) {
    if (params.get_value<bool>("PhotonSourceDistribution:output sources", false))
      cmi_error("PhotonSourceDistribution:output sources is not provided by the B200 backend!");
  }
  size_t get_number_of_sources() const override { return num_sources_; }
  Vec3 get_position(size_t index) override {
    if (index > num_sources_) cmi_error("Source index out of range!");
    Vec3 p;
    p[0] = anchor_x_ + random_generator_.get_uniform_random_double() * sides_x_;
    p[1] = anchor_y_ + random_generator_.get_uniform_random_double() * sides_y_;
    const double rho = scaleheight_z_ * std::sqrt(-2. * std::log(random_generator_.get_uniform_random_double()));
    p[2] = rho * std::cos(2. * M_PI * random_generator_.get_uniform_random_double()) + origin_z_;
    return p;
  }
  double get_weight(size_t) const override { return 1. / num_sources_; }
  double get_total_luminosity() const override { return num_sources_ * luminosity_; }

private:
  uint32_t num_sources_;
  double anchor_x_, sides_x_, anchor_y_, sides_y_, origin_z_, scaleheight_z_, luminosity_;
  RandomGenerator random_generator_;
};

/* Sources from an SPH snapshot (GadgetSnapshotPhotonSourceDistribution.cpp:60-325): the star particles of
 * /PartType4 inside the simulation box (or, with `use gas`, the star-forming gas particles of /PartType0 with a
 * stellar mass SFR x cutoff age), each with the UV luminosity of its age and mass.  UVLuminosityFunction:
 * RateBased (RateBasedUVLuminosityFunction.hpp: mass x rate while younger than the cutoff age, the factory's
 * default); IMFBased needs the stellar-population sampling of the RHD drivers and is refused.  Read with
 * host/HDF5Reader.hpp. */
class GadgetSnapshotPhotonSourceDistribution : public PhotonSourceDistribution {
public:
  explicit GadgetSnapshotPhotonSourceDistribution(ParameterFile &params, Log *log = nullptr) {
    const std::string filename = params.get_filename("PhotonSourceDistribution:filename");
    const std::string formation_time_name =
        params.get_value<std::string>("PhotonSourceDistribution:formation time name", "FormationTime");
    const Vec3 anchor = params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:anchor");
    const Vec3 sides = params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:sides");
    const std::string lf_type = params.get_value<std::string>("UVLuminosityFunction:type", "RateBased");
    if (lf_type != "RateBased")
      cmi_error("Unknown UVLuminosityFunction type: \"%s\" (the B200 backend provides RateBased).", lf_type.c_str());
    const double UV_rate_per_mass_unit =
        params.get_physical_value<QUANTITY_FREQUENCY_PER_MASS>("UVLuminosityFunction:UV rate per mass unit", "2.49428e16 s^-1 kg^-1");
    const double lf_cutoff_age = params.get_physical_value<QUANTITY_TIME>("UVLuminosityFunction:cutoff age", "5. Myr");
    const double fallback_unit_length_in_SI = params.get_physical_value<QUANTITY_LENGTH>("PhotonSourceDistribution:fallback unit length", "0. m");
    const double fallback_unit_time_in_SI = params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:fallback unit time", "0. s");
    const double fallback_unit_mass_in_SI = params.get_physical_value<QUANTITY_MASS>("PhotonSourceDistribution:fallback unit mass", "0. kg");
    const double cutoff_age = params.get_physical_value<QUANTITY_TIME>("PhotonSourceDistribution:cutoff age", "5. Myr");
    const bool use_gas = params.get_value<bool>("PhotonSourceDistribution:use gas", false);
    const double SFR_unit = params.get_physical_value<QUANTITY_MASS_RATE>("PhotonSourceDistribution:SFR unit", "0. kg s^-1");
    const bool comoving_integration = params.get_value<bool>("PhotonSourceDistribution:comoving integration flag", false);
    const double hubble_parameter = params.get_value<double>("PhotonSourceDistribution:hubble parameter", 0.7);
    auto luminosity_function = [&](double age, double mass) { return age <= lf_cutoff_age ? mass * UV_rate_per_mass_unit : 0.; };

    hdf5::HDF5Input file(filename);
    const double snaptime = file.read_double_attribute("/Header", "Time")[0];
    double unit_length_in_SI = fallback_unit_length_in_SI, unit_time_in_SI = fallback_unit_time_in_SI,
           unit_mass_in_SI = fallback_unit_mass_in_SI;
    if (file.exists("/Units")) {
      unit_length_in_SI = UnitConverter::to_SI(QUANTITY_LENGTH, file.read_double_attribute("/Units", "Unit length in cgs (U_L)")[0], "cm");
      unit_time_in_SI = file.read_double_attribute("/Units", "Unit time in cgs (U_t)")[0];
      unit_mass_in_SI = UnitConverter::to_SI(QUANTITY_MASS, file.read_double_attribute("/Units", "Unit mass in cgs (U_M)")[0], "g");
    } else {
      if (log) log->write_warning("No Units group found! Using fallback units.");
      if (unit_length_in_SI == 0.) unit_length_in_SI = 1.;
      if (unit_time_in_SI == 0.) unit_time_in_SI = 1.;
      if (unit_mass_in_SI == 0.) unit_mass_in_SI = 1.;
    }
    if (comoving_integration) {
      unit_length_in_SI /= hubble_parameter;
      unit_mass_in_SI /= hubble_parameter;
      unit_time_in_SI /= hubble_parameter;
    }
    auto inside = [&](const Vec3 &v) { /* Box::inside (Box.hpp:191-195) */
      return v[0] >= anchor[0] && v[0] < anchor[0] + sides[0] && v[1] >= anchor[1] && v[1] < anchor[1] + sides[1] &&
             v[2] >= anchor[2] && v[2] < anchor[2] + sides[2];
    };
    total_luminosity_ = 0.;
    const std::string group = use_gas ? "/PartType0" : "/PartType4";
    std::vector<uint64_t> dims;
    const std::vector<double> x = file.read_dataset(group + "/Coordinates", &dims);
    if (dims.size() != 2 || dims[1] != 3) cmi_error("Snapshot \"%s\": bad %s/Coordinates!", filename.c_str(), group.c_str());
    const size_t n = dims[0];
    std::vector<double> a, b;
    if (use_gas) {
      a = file.read_dataset("/PartType0/StarFormationRate");
    } else {
      a = file.read_dataset("/PartType4/" + formation_time_name);
      b = file.read_dataset("/PartType4/Masses");
    }
    if (a.size() != n || (!use_gas && b.size() != n)) cmi_error("Snapshot \"%s\": datasets of %s differ in length!", filename.c_str(), group.c_str());
    const double unit_SFR_in_SI = (SFR_unit == 0.) ? unit_mass_in_SI / unit_time_in_SI : SFR_unit;
    for (size_t i = 0; i < n; ++i) {
      const Vec3 position = {x[3 * i] * unit_length_in_SI, x[3 * i + 1] * unit_length_in_SI, x[3 * i + 2] * unit_length_in_SI};
      double UV_luminosity = 0.;
      if (use_gas) {
        if (a[i] > 0. && inside(position)) UV_luminosity = luminosity_function(0., a[i] * unit_SFR_in_SI * cutoff_age);
      } else if (inside(position)) {
        UV_luminosity = luminosity_function((snaptime - a[i]) * unit_time_in_SI, b[i] * unit_mass_in_SI);
      }
      if (UV_luminosity > 0.) {
        positions_.push_back(position);
        luminosities_.push_back(UV_luminosity);
        total_luminosity_ += UV_luminosity;
      }
    }
    if (log) log->write_status("Found ", positions_.size(), " active sources, with a total luminosity of ", total_luminosity_, " s^-1.");
  }
  size_t get_number_of_sources() const override { return positions_.size(); }
  Vec3 get_position(size_t i) override { return positions_[i]; }
  double get_weight(size_t i) const override { return luminosities_[i] / total_luminosity_; }
  double get_total_luminosity() const override { return total_luminosity_; }

private:
  std::vector<Vec3> positions_;
  std::vector<double> luminosities_;
  double total_luminosity_ = 0.;
};

struct PhotonSourceDistributionFactory {
  static PhotonSourceDistribution *generate(ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("PhotonSourceDistribution:type", "SingleStar");
    if (log) log->write_info("Requested PhotonSourceDistribution type: ", type);
    if (type == "SingleStar") return new SingleStarPhotonSourceDistribution(params);
    if (type == "AsciiFile") return new AsciiFilePhotonSourceDistribution(params);
    if (type == "AsciiFileTable") return new AsciiFileTablePhotonSourceDistribution(params);
    if (type == "UniformRandom") return new UniformRandomPhotonSourceDistribution(params);
    if (type == "DiscPatch") return new DiscPatchPhotonSourceDistribution(params);
    if (type == "DwarfGalaxy") return new DwarfGalaxyPhotonSourceDistribution(params);
    if (type == "SILCC") return new SILCCPhotonSourceDistribution(params);
    if (type == "GadgetSnapshot") return new GadgetSnapshotPhotonSourceDistribution(params, log);
    if (type == "None") return nullptr;
    cmi_error("Unknown PhotonSourceDistribution type: \"%s\" (the B200 backend provides SingleStar, AsciiFile, "
              "AsciiFileTable, UniformRandom, DiscPatch, DwarfGalaxy, SILCC and GadgetSnapshot)!",
              type.c_str());
  }
};

/* ---- plugins that are pure parameters for the device ---- */
struct PhotonSourceSpectrum {
  int kind;     /* CMIB_SPECTRUM_* */
  double param; /* frequency (Hz) or temperature (K) */
  double total_flux = -1.;
  /* CMIB_SPECTRUM_TABULATED: the two arrays the device samples from (cmib_set_spectrum_table) */
  std::vector<double> frequencies, cumulative_distribution;

  /* hand the spectrum to a device context: role 0 = discrete sources, 1 = continuous source */
  int set_on(cmib_context *ctx, int role) const {
    if (kind == CMIB_SPECTRUM_TABULATED)
      return cmib_set_spectrum_table(ctx, role, (int32_t)frequencies.size(), frequencies.data(),
                                     cumulative_distribution.data());
    return role == 0 ? cmib_set_spectrum(ctx, kind, param) : 0; /* role 1: through cmib_set_continuous_source */
  }

  /* get_random_frequency on the host, with the reference's generator: the same functions the device
   * runs (csrc/source.cuh), used to build a Masked spectrum */
  double sample(RandomGenerator &random_generator) {
    if (kind == CMIB_SPECTRUM_MONOCHROMATIC) return param;
    const double x = random_generator.get_uniform_random_double();
    if (kind == CMIB_SPECTRUM_PLANCK) {
      if (planck_table_.empty()) cmib::host::build_planck_table(param, planck_table_);
      return cmib::planck_frequency_at(planck_table_.data(), x);
    }
    if (kind == CMIB_SPECTRUM_UNIFORM) return cmib::uniform_frequency(x);
    return cmib::tabulated_frequency(frequencies.data(), cumulative_distribution.data(), (uint32_t)frequencies.size(), x);
  }
  std::vector<double> planck_table_;

  /*
   * MaskedPhotonSourceSpectrum (src/MaskedPhotonSourceSpectrum.cpp:40-122): another spectrum seen through
   * a frequency-dependent mask.  The unmasked spectrum is sampled `mask number of samples` times with
   * RandomGenerator() (seed 42) into `mask number of bins` bins between 13.6 and 54.4 eV, every bin is
   * multiplied by the mask (Linear: 1 at 13.6 eV falling to 0 at 54.4 eV,
   * LinearPhotonSourceSpectrumMask.hpp:42-49), the result is made cumulative and normalised: a tabulated
   * spectrum for the device.  Same generator, same samplers: the table is the reference's bit for bit.
   */
  static PhotonSourceSpectrum *masked(const std::string &role, ParameterFile &params, Log *log) {
    const std::string unmasked_type = params.get_value<std::string>(role + ":masked type", "Planck");
    if (unmasked_type == "Masked") cmi_error("A Masked spectrum cannot mask itself!");
    std::unique_ptr<PhotonSourceSpectrum> unmasked(generate_from_type(unmasked_type, role, params, log));
    if (!unmasked) cmi_error("No spectrum to mask!");
    const std::string mask_type = params.get_value<std::string>(role + ":PhotonSourceSpectrumMask:type", "Linear");
    if (mask_type != "Linear") cmi_error("Unknown PhotonSourceSpectrumMask type: \"%s\"!", mask_type.c_str());
    const uint32_t number_of_bins = params.get_value<uint32_t>(role + ":mask number of bins", 1000);
    const uint32_t number_of_samples = params.get_value<uint32_t>(role + ":mask number of samples", 10000000);
    auto *s = new PhotonSourceSpectrum{CMIB_SPECTRUM_TABULATED, 0.};
    std::vector<double> &freq = s->frequencies, &cdf = s->cumulative_distribution;
    freq.assign(number_of_bins, 0.);
    cdf.assign(number_of_bins, 0.);
    const double min_frequency = 3.289e15, max_frequency = 4. * min_frequency;
    const double frequency_bin_size = (max_frequency - min_frequency) / (number_of_bins - 1.);
    for (uint32_t i = 0; i < number_of_bins; ++i) freq[i] = min_frequency + i * frequency_bin_size;
    RandomGenerator random_generator;
    for (uint32_t i = 0; i < number_of_samples; ++i) {
      const double random_frequency = unmasked->sample(random_generator);
      const uint32_t index = (uint32_t)((random_frequency - min_frequency) / frequency_bin_size);
      if (index < number_of_bins) cdf[index] += 1.; /* the reference writes out of bounds otherwise */
    }
    for (uint32_t i = 0; i < number_of_bins; ++i) cdf[i] *= 1. - (freq[i] - min_frequency) / (max_frequency - min_frequency);
    for (uint32_t i = 1; i < number_of_bins; ++i) cdf[i] += cdf[i - 1];
    const double norm = cdf.back();
    const double norm_inv = 1. / norm;
    for (uint32_t i = 0; i < number_of_bins; ++i) cdf[i] *= norm_inv;
    s->total_flux = norm * unmasked->total_flux / number_of_samples;
    return s;
  }

  /* Utilities::locate (src/Utilities.hpp:726-742) */
  static uint32_t locate(double x, const double *xarr, uint32_t length) {
    uint32_t jl = 0, ju = length;
    while (ju - jl > 1) {
      const uint32_t jm = (ju + jl) >> 1;
      if (x > xarr[jm]) jl = jm; else ju = jm;
    }
    if (jl == length - 1) --jl;
    return jl;
  }

  /*
   * FaucherGiguerePhotonSourceSpectrum (src/FaucherGiguerePhotonSourceSpectrum.cpp:40-183): the UV
   * background of Faucher-Giguere et al. (2009, December 2011 tables) at a redshift, resampled on 100
   * frequencies between 13.6 and 54.4 eV.  Data files: <CMIB_DATA_DIR>/fg_uvb_dec11/ (the unpacked
   * data/fg_uvb_dec11.tar.gz of a CMacIonize checkout; the reference's build unpacks it likewise).
   * NB the reference reads the second redshift table from the FIRST file's (exhausted) stream
   * (:96-104), which leaves the added term undefined; it is multiplied by zero when the redshift is
   * a multiple of 0.05, where this function is bit-identical (tests/test_host_layer.py).  In between
   * it interpolates the two tables as the reference's comments say it intends to.
   */
  static PhotonSourceSpectrum *faucher_giguere(double redshift) {
    constexpr int NUMFREQ = 100;
    auto *s = new PhotonSourceSpectrum{CMIB_SPECTRUM_TABULATED, redshift};
    s->frequencies.assign(NUMFREQ, 0.);
    s->cumulative_distribution.assign(NUMFREQ, 0.);
    std::vector<double> &freq = s->frequencies, &cdf = s->cumulative_distribution;
    const double min_frequency = 3.289e15, max_frequency = 4. * min_frequency;
    for (int i = 0; i < NUMFREQ; ++i) freq[i] = min_frequency + i * (max_frequency - min_frequency) / (NUMFREQ - 1.);
    s->total_flux = 0.;
    if (!(redshift <= 10.65)) return s; /* no UV background: all zeros, like the reference */
    const char *dir = getenv("CMIB_DATA_DIR");
    if (!dir) cmi_error("FaucherGiguere spectrum: set CMIB_DATA_DIR to the directory that holds fg_uvb_dec11/!");
    auto filename = [&](double z) { /* get_filename (:194-216): integer arithmetic on z / 0.05 */
      uint32_t iz = (uint32_t)(std::round(z / 0.05) * 5);
      const uint32_t iz100 = iz / 100;
      iz -= iz100 * 100;
      const uint32_t iz10 = iz / 10;
      iz -= iz10 * 10;
      std::ostringstream name;
      name << dir << "/fg_uvb_dec11/fg_uvb_dec11_z_" << iz100 << "." << iz10;
      if (iz > 0) name << iz;
      name << ".dat";
      return name.str();
    };
    auto read = [&](double z, double fac, double *nu_out, double *ener, bool add) {
      const std::string name = filename(z);
      std::ifstream file(name);
      if (!file) cmi_error("File not found: %s!", name.c_str());
      std::string line;
      getline(file, line);
      getline(file, line);
      for (int i = 0; i < 261; ++i) {
        getline(file, line);
        std::istringstream linestream(line);
        double nu = 0., e = 0.;
        linestream >> nu >> e;
        if (nu_out) nu_out[i] = nu * 3.289e15;
        if (add) ener[i] += fac * e; else ener[i] = fac * e;
      }
    };
    double spectrum_freq[261], spectrum_ener[261];
    const unsigned int izlo = (unsigned int)(redshift / 0.05);
    const unsigned int izhi = izlo + 1;
    const double zlo = izlo * 0.05, zhi = izhi * 0.05;
    read(zlo, 20. * (zhi - redshift), spectrum_freq, spectrum_ener, false);
    const double zhi_fac = 20. * (redshift - zlo);
    if (zhi <= 10.65 && zhi_fac != 0.) read(zhi, zhi_fac, nullptr, spectrum_ener, true);
    for (int i = 1; i < NUMFREQ; ++i) {
      const double y1 = freq[i - 1];
      const uint32_t i1 = locate(y1, spectrum_freq, 261);
      double f = (y1 - spectrum_freq[i1]) / (spectrum_freq[i1 + 1] - spectrum_freq[i1]);
      const double e1 = spectrum_ener[i1] + f * (spectrum_ener[i1 + 1] - spectrum_ener[i1]);
      const double y2 = freq[i];
      const uint32_t i2 = locate(y2, spectrum_freq, 261);
      f = (y2 - spectrum_freq[i2]) / (spectrum_freq[i2 + 1] - spectrum_freq[i2]);
      const double e2 = spectrum_ener[i2] + f * (spectrum_ener[i2 + 1] - spectrum_ener[i2]);
      cdf[i] = 0.5 * (e1 / y2 + e2 / y1) * (y2 - y1);
    }
    for (int i = 1; i < NUMFREQ; ++i) cdf[i] += cdf[i - 1];
    /* 1e-21 erg Hz^-1 s^-1 cm^-2 sr^-1 -> m^-2 s^-1 (:150-161) */
    s->total_flux = 1.e-28 * cdf[NUMFREQ - 1] / 6.626070040e-34;
    s->total_flux *= 4. * M_PI;
    s->total_flux *= 1.e4;
    const double norm = cdf[NUMFREQ - 1];
    for (int i = 0; i < NUMFREQ; ++i) cdf[i] /= norm;
    return s;
  }
  static PhotonSourceSpectrum *generate(const std::string &role, ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>(role + ":type", "Monochromatic");
    if (log) log->write_info("Requested PhotonSourceSpectrum for ", role, ": ", type);
    return generate_from_type(type, role, params, log);
  }
  /* PhotonSourceSpectrumFactory::generate_from_type (src/PhotonSourceSpectrumFactory.hpp:84-119) */
  static PhotonSourceSpectrum *generate_from_type(const std::string &type, const std::string &role, ParameterFile &params,
                                                  Log *log = nullptr) {
    if (type == "Masked") return masked(role, params, log);
    if (type == "Monochromatic") {
      auto *s = new PhotonSourceSpectrum{CMIB_SPECTRUM_MONOCHROMATIC,
                                         params.get_physical_value<QUANTITY_FREQUENCY>(role + ":frequency", "13.6 eV")};
      s->total_flux = params.get_physical_value<QUANTITY_FLUX>(role + ":total flux", "-1. m^-2 s^-1");
      return s;
    }
    if (type == "Planck") {
      auto *s = new PhotonSourceSpectrum{CMIB_SPECTRUM_PLANCK,
                                         params.get_physical_value<QUANTITY_TEMPERATURE>(role + ":temperature", "4.e4 K")};
      s->total_flux = params.get_physical_value<QUANTITY_FLUX>(role + ":ionizing flux", "-1. m^-2 s^-1");
      return s;
    }
    if (type == "Uniform") return new PhotonSourceSpectrum{CMIB_SPECTRUM_UNIFORM, 0.}; /* no total flux (UniformPhotonSourceSpectrum.hpp:60-63) */
    if (type == "FaucherGiguere") return faucher_giguere(params.get_value<double>(role + ":redshift", 0.));
    if (type == "None") return nullptr;
    cmi_error("Unknown PhotonSourceSpectrum type: \"%s\" (the B200 backend provides Monochromatic, Planck, Uniform, "
              "FaucherGiguere and Masked; any tabulated spectrum can be handed to cmib_set_spectrum_table)!",
              type.c_str());
  }
};

struct CrossSections {
  int kind; /* CMIB_CROSS_SECTIONS_*; 2 = Bimodal (two constant values per ion) */
  double fixed[CMIB_NUM_IONS] = {0.};
  double high[CMIB_NUM_IONS] = {0.};
  double frequency_limit = 0.;
  int set_on(cmib_context *ctx) const {
    if (kind == 2) return cmib_set_bimodal_cross_sections(ctx, frequency_limit, fixed, high);
    return cmib_set_cross_sections(ctx, kind, fixed);
  }
  static CrossSections *generate(ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("CrossSections:type", "Verner");
    if (log) log->write_info("Requested CrossSections type: ", type);
    auto *c = new CrossSections();
    if (type == "Verner") {
      c->kind = CMIB_CROSS_SECTIONS_VERNER;
    } else if (type == "FixedValue") {
      c->kind = CMIB_CROSS_SECTIONS_FIXED_VALUE;
      static const char *keys[CMIB_NUM_IONS] = {"hydrogen_0", "helium_0", "carbon_1", "carbon_2", "nitrogen_0",
                                                "nitrogen_1", "nitrogen_2", "oxygen_0", "oxygen_1", "neon_0",
                                                "neon_1", "sulphur_1", "sulphur_2", "sulphur_3"};
      for (int i = 0; i < CMIB_NUM_IONS; ++i)
        c->fixed[i] = params.get_physical_value<QUANTITY_SURFACE_AREA>(std::string("CrossSections:") + keys[i],
                                                                      i == 0 ? "6.3e-18 cm^2" : "0. m^2");
    } else if (type == "Bimodal") {
      /* BiModalCrossSections(ParameterFile&) (src/BimodalCrossSections.hpp:174-245), kept as it is: the
       * frequency limit is read from the key "frequency limit:" (no block), and the member initialisers
       * swap the two values of oxygen_0 and of sulphur_1 (:132, :138 / :151, :157): "oxygen_0_high" is
       * what applies BELOW the limit */
      c->kind = 2;
      c->frequency_limit = params.get_physical_value<QUANTITY_FREQUENCY>("frequency limit:", "15. eV");
      static const char *keys[CMIB_NUM_IONS] = {"hydrogen_0", "helium_0", "carbon_1", "carbon_2", "nitrogen_0",
                                                "nitrogen_1", "nitrogen_2", "oxygen_0", "oxygen_1", "neon_0",
                                                "neon_1", "sulphur_1", "sulphur_2", "sulphur_3"};
      for (int i = 0; i < CMIB_NUM_IONS; ++i) {
        const double low = params.get_physical_value<QUANTITY_SURFACE_AREA>(std::string("CrossSections:") + keys[i] + "_low",
                                                                            i == 0 ? "6.3e-18 cm^2" : "0. m^2");
        const double high = params.get_physical_value<QUANTITY_SURFACE_AREA>(std::string("CrossSections:") + keys[i] + "_high",
                                                                             i == 0 ? "6.3e-18 cm^2" : "0. m^2");
        const bool swapped = (i == 7 || i == 11); /* oxygen_0, sulphur_1 */
        c->fixed[i] = swapped ? high : low;
        c->high[i] = swapped ? low : high;
      }
    } else {
      delete c;
      cmi_error("Unknown CrossSections type: \"%s\"!", type.c_str());
    }
    return c;
  }
};

struct RecombinationRates {
  int kind;
  double fixed[CMIB_NUM_IONS] = {0.};
  static RecombinationRates *generate(ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("RecombinationRates:type", "Verner");
    if (log) log->write_info("Requested RecombinationRates type: ", type);
    auto *r = new RecombinationRates();
    if (type == "Verner") {
      r->kind = CMIB_RECOMBINATION_VERNER;
    } else if (type == "FixedValue") {
      r->kind = CMIB_RECOMBINATION_FIXED_VALUE;
      static const char *keys[CMIB_NUM_IONS] = {"hydrogen_1", "helium_1", "carbon_2", "carbon_3", "nitrogen_1",
                                                "nitrogen_2", "nitrogen_3", "oxygen_1", "oxygen_2", "neon_1",
                                                "neon_2", "sulphur_2", "sulphur_3", "sulphur_4"};
      for (int i = 0; i < CMIB_NUM_IONS; ++i)
        r->fixed[i] = params.get_physical_value<QUANTITY_REACTION_RATE>(
            std::string("RecombinationRates:") + keys[i], i == 0 ? "2.7e-13 cm^3 s^-1" : "0. m^3 s^-1");
    } else {
      delete r;
      cmi_error("Unknown RecombinationRates type: \"%s\"!", type.c_str());
    }
    return r;
  }
};

struct Abundances {
  double abundance[CMIB_NUM_ELEMENTS] = {0.};
  static Abundances generate(ParameterFile &params, Log *log = nullptr) {
    /* deprecated "Abundances:helium" style block -> AbundanceModel (AbundanceModelFactory.hpp:54-86) */
    static const char *old_names[CMIB_NUM_ELEMENTS] = {"helium", "carbon", "nitrogen", "oxygen", "neon", "sulphur"};
    if (!params.has_value("AbundanceModel:type")) {
      bool migrated = false;
      for (int i = 0; i < CMIB_NUM_ELEMENTS; ++i) {
        const std::string old_key = std::string("Abundances:") + old_names[i];
        if (params.has_value(old_key)) {
          params.add_value(std::string("AbundanceModel:") + element_name(i), params.get_value<std::string>(old_key));
          migrated = true;
        }
      }
      if (migrated) {
        params.add_value("AbundanceModel:type", "FixedValue");
        if (log) log->write_warning("Deprecated Abundances block converted to AbundanceModel:type FixedValue.");
      }
    }
    const std::string type = params.get_value<std::string>("AbundanceModel:type", "FixedValue");
    Abundances a;
    if (type == "SolarMetallicity") {
      /* SolarMetallicityAbundanceModel (src/SolarMetallicityAbundanceModel.hpp:46-121): log10 abundances
       * scaled with the oxygen abundance (N with its secondary-production break at -4) */
      const double metallicity = params.get_value<double>("AbundanceModel:metallicity", -3.31);
      const double solar_He = -1.07, solar_C = -3.57, solar_N = -4.17, solar_O = -3.31, solar_Ne = -4.07, solar_S = -4.88;
      double actual_C = solar_C, actual_N = solar_N, actual_Ne = solar_Ne, actual_S = solar_S;
      if (metallicity != solar_O) {
        const double Odiff = metallicity - solar_O;
        actual_C = solar_C + Odiff;
        actual_Ne = solar_Ne + Odiff;
        actual_S = solar_S + Odiff;
        actual_N = (metallicity <= -4.) ? metallicity - 1.6 : metallicity + 0.6 * (metallicity + 4.) - 1.6;
      }
      const double logs[CMIB_NUM_ELEMENTS] = {solar_He, actual_C, actual_N, metallicity, actual_Ne, actual_S};
      for (int i = 0; i < CMIB_NUM_ELEMENTS; ++i) a.abundance[i] = std::pow(10., logs[i]);
      return a;
    }
    if (type != "FixedValue") cmi_error("Unknown AbundanceModel type: \"%s\"!", type.c_str());
    for (int i = 0; i < CMIB_NUM_ELEMENTS; ++i)
      a.abundance[i] = params.get_value<double>(std::string("AbundanceModel:") + element_name(i), 0.);
    return a;
  }
};

struct DiffuseReemissionHandler {
  int kind = CMIB_REEMISSION_NONE;
  double probability = 0.364, frequency = 0.;
  static DiffuseReemissionHandler generate(ParameterFile &params, Log *log = nullptr) {
    if (!params.has_value("DiffuseReemissionHandler:type") && params.has_value("PhotonSource:diffuse field")) {
      if (log) log->write_warning("\"PhotonSource:diffuse field\" was replaced by \"DiffuseReemissionHandler\"; converting.");
      const bool on = params.get_value<bool>("PhotonSource:diffuse field", false);
      params.add_value("DiffuseReemissionHandler:type", on ? "Physical" : "None");
    }
    const std::string type = params.get_value<std::string>("DiffuseReemissionHandler:type", "None");
    if (log) log->write_info("Requested DiffuseReemissionHandler type: ", type);
    DiffuseReemissionHandler h;
    if (type == "FixedValue") {
      h.kind = CMIB_REEMISSION_FIXED_VALUE;
      h.probability = params.get_value<double>("DiffuseReemissionHandler:reemission probability", 0.364);
      h.frequency = params.get_physical_value<QUANTITY_FREQUENCY>("DiffuseReemissionHandler:reemission frequency", "19.8 eV");
    } else if (type == "Physical") {
      h.kind = CMIB_REEMISSION_PHYSICAL;
    } else if (type == "None") {
      h.kind = CMIB_REEMISSION_NONE;
    } else {
      cmi_error("Unknown DiffuseReemissionHandler type: \"%s\"!", type.c_str());
    }
    return h;
  }
};

inline cmib_temperature_params temperature_calculator_parameters(ParameterFile &params) {
  cmib_temperature_params p;
  p.do_temperature_calculation = params.get_value<bool>("TemperatureCalculator:do temperature calculation", false);
  p.minimum_number_of_iterations = params.get_value<uint32_t>("TemperatureCalculator:minimum number of iterations", 3);
  p.epsilon_convergence = params.get_value<double>("TemperatureCalculator:epsilon convergence", 1.e-3);
  p.maximum_number_of_iterations = params.get_value<uint32_t>("TemperatureCalculator:maximum number of iterations", 100);
  p.pah_heating_factor = params.get_value<double>("TemperatureCalculator:PAH heating factor", 0.);
  p.cosmic_ray_heating_factor = params.get_value<double>("TemperatureCalculator:cosmic ray heating factor", 0.);
  p.cosmic_ray_heating_limit = params.get_value<double>("TemperatureCalculator:cosmic ray heating limit", 0.75);
  p.cosmic_ray_heating_scale_length =
      params.get_physical_value<QUANTITY_LENGTH>("TemperatureCalculator:cosmic ray heating scale length", "1.33333 kpc");
  p.minimum_ionized_temperature =
      params.get_physical_value<QUANTITY_TEMPERATURE>("TemperatureCalculator:minimum ionized temperature", "4000. K");
  return p;
}

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

/* ---- writers ---- */
class DensityGridWriter {
public:
  virtual ~DensityGridWriter() {}
  /* DensityGridWriter::write(grid, iteration, params, time) (DensityGridWriter.hpp); works on the host mirror
   * of the cells: the caller refreshes it (CartesianDensityGrid::download) */
  virtual void write(CartesianCells &grid, uint32_t iteration, ParameterFile &params, double time = 0.) = 0;
};

/* the reference's ASCII snapshot layout, optionally with every field */
class AsciiFileDensityGridWriter : public DensityGridWriter {
public:
  AsciiFileDensityGridWriter(std::string prefix, std::string output_folder, bool all_fields = false)
      : prefix_(std::move(prefix)), folder_(std::move(output_folder)), all_fields_(all_fields) {}
  AsciiFileDensityGridWriter(const std::string &output_folder, ParameterFile &params)
      : AsciiFileDensityGridWriter(params.get_value<std::string>("DensityGridWriter:prefix", "snapshot"), output_folder,
                                   params.get_value<bool>("DensityGridWriter:all fields", false)) {}
  std::string filename(uint32_t iteration) const {
    char num[16];
    snprintf(num, sizeof(num), "%03u", iteration);
    return folder_ + "/" + prefix_ + num + ".txt";
  }
  void write(CartesianCells &grid, uint32_t iteration, ParameterFile &, double = 0.) override { write(grid, iteration); }
  void write(CartesianCells &grid, uint32_t iteration) {
    std::ofstream file(filename(iteration));
    if (!file) cmi_error("Unable to open snapshot file \"%s\"!", filename(iteration).c_str());
    const size_t n = grid.get_number_of_cells();
    const double volume = grid.get_cell_volume();
    if (!all_fields_) {
      /* AsciiFileDensityGridWriter.cpp:75-95 */
      file << "#x (m)\ty (m)\tz (m)\tn (m^-3)\tvolume (m^3)\tneutral H fraction\n";
      for (size_t i = 0; i < n; ++i) {
        const Vec3 x = grid.get_cell_midpoint(i);
        file << x[0] << "\t" << x[1] << "\t" << x[2] << "\t" << grid.number_density[i] << "\t" << volume << "\t"
             << grid.ionic_fraction[i] << "\n";
      }
    } else {
      file << "#x (m)\ty (m)\tz (m)\tn (m^-3)\tvolume (m^3)\tT (K)";
      for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) file << "\tNeutralFraction" << ion_name(ion);
      file << "\n" << std::setprecision(17);
      for (size_t i = 0; i < n; ++i) {
        const Vec3 x = grid.get_cell_midpoint(i);
        file << x[0] << "\t" << x[1] << "\t" << x[2] << "\t" << grid.number_density[i] << "\t" << volume << "\t"
             << grid.temperature[i];
        for (int ion = 0; ion < CMIB_NUM_IONS; ++ion) file << "\t" << grid.ionic_fraction[(size_t)ion * n + i];
        file << "\n";
      }
    }
  }

private:
  std::string prefix_, folder_;
  bool all_fields_;
};

/* Which cell properties a snapshot holds: the `DensityGridWriterFields:` block
 * (DensityGridWriterFields.hpp:790-835).  Without hydro the defaults are Coordinates, NumberDensity and
 * NeutralFractionH; every `NeutralFraction<ion>` and `Temperature` can be switched on.  As in the reference a
 * flagged ion also switches on the ions before it (`ion_present` shifts the flag word, :843-846). */
struct DensityGridWriterFields {
  bool coordinates, number_density, temperature;
  uint32_t neutral_fraction = 0;
  explicit DensityGridWriterFields(ParameterFile &params) {
    coordinates = params.get_value<uint32_t>("DensityGridWriterFields:Coordinates", 1) > 0;
    number_density = params.get_value<uint32_t>("DensityGridWriterFields:NumberDensity", 1) > 0;
    temperature = params.get_value<uint32_t>("DensityGridWriterFields:Temperature", 0) > 0;
    for (int ion = 0; ion < CMIB_NUM_IONS; ++ion)
      neutral_fraction += params.get_value<uint32_t>(std::string("DensityGridWriterFields:NeutralFraction") + ion_symbol(ion),
                                                     ion == 0 ? 1u : 0u)
                          << ion;
    if (params.get_value<uint32_t>("DensityGridWriterFields:CosmicRayFactor", 0) > 0)
      cmi_error("DensityGridWriterFields:CosmicRayFactor is not provided by the B200 backend!");
  }
  bool ion_present(int ion) const { return (neutral_fraction >> ion) > 0; }
};

/* Gadget-style HDF5 snapshot, group for group and attribute for attribute what GadgetDensityGridWriter::write
 * produces (GadgetDensityGridWriter.cpp:122-300): /Header, /Code, /Configuration, /Parameters (the used values),
 * /RuntimePars, /Units (SI) and /PartType0 with Coordinates (relative to the box anchor), NumberDensity,
 * Temperature and NeutralFraction<ion>, so that the reference's benchmark analysis scripts read it
 * unchanged.  Written by host/HDF5Writer.hpp; datasets are contiguous, never compressed. */
class GadgetDensityGridWriter : public DensityGridWriter {
public:
  GadgetDensityGridWriter(std::string prefix, std::string output_folder, const DensityGridWriterFields &fields,
                          uint32_t padding = 3)
      : prefix_(std::move(prefix)), folder_(std::move(output_folder)), fields_(fields), padding_(padding) {}
  GadgetDensityGridWriter(const std::string &output_folder, ParameterFile &params)
      : GadgetDensityGridWriter(params.get_value<std::string>("DensityGridWriter:prefix", "snapshot"), output_folder,
                                DensityGridWriterFields(params), params.get_value<uint32_t>("DensityGridWriter:padding", 3)) {
    if (params.get_value<bool>("DensityGridWriter:compression", false))
      cmi_error("DensityGridWriter:compression is not provided by the B200 backend!");
  }
  /* Utilities::compose_filename: folder/prefixNNN.hdf5 */
  std::string filename(uint32_t iteration) const {
    char num[32];
    snprintf(num, sizeof(num), "%0*u", (int)padding_, iteration);
    return folder_ + "/" + prefix_ + num + ".hdf5";
  }
  void write(CartesianCells &grid, uint32_t iteration, ParameterFile &params, double time = 0.) override {
    const size_t n = grid.get_number_of_cells();
    hdf5::HDF5File file;
    hdf5::Group &header = file.root().create_group("Header");
    header.write_attribute("BoxSize", grid.get_box_sides());
    header.write_attribute("Dimension", int32_t(3));
    header.write_attribute("Flag_Entropy_ICs", std::vector<uint32_t>(6, 0));
    header.write_attribute("MassTable", std::vector<double>(6, 0.));
    header.write_attribute("NumFilesPerSnapshot", int32_t(1));
    std::vector<uint32_t> numpart(6, 0);
    numpart[0] = (uint32_t)n;
    header.write_attribute("NumPart_ThisFile", numpart);
    header.write_attribute("NumPart_Total", numpart);
    header.write_attribute("NumPart_Total_HighWord", std::vector<uint32_t>(6, 0));
    header.write_attribute("Time", time);

    hdf5::Group &code = file.root().create_group("Code");
    struct utsname os;
    if (uname(&os) != 0) memset(&os, 0, sizeof(os));
    code.write_attribute("Git version", "cmacionize_b200 (C ABI " + std::to_string(cmib_abi_version()) + ")");
    code.write_attribute("Compilation date", __DATE__);
    code.write_attribute("Compilation time", __TIME__);
    code.write_attribute("Compiler", std::string("GNU ") + __VERSION__);
    code.write_attribute("Operating system", os.sysname);
    code.write_attribute("Kernel name", std::string(os.sysname) + " " + os.release);
    code.write_attribute("Hardware name", os.machine);
    code.write_attribute("Host name", os.nodename);

    hdf5::Group &configuration = file.root().create_group("Configuration");
    configuration.write_attribute("BACKEND", "B200 (sm_100a) photoionization hot path, libcmib.so");
    configuration.write_attribute("HAVE_HDF5", "False (built-in writer: host/HDF5Writer.hpp)");
    configuration.write_attribute("NUMBER_OF_IONNAMES", std::to_string(CMIB_NUM_IONS));

    hdf5::Group &parameters = file.root().create_group("Parameters");
    for (const auto &kv : params.used_values()) parameters.write_attribute(kv.first, kv.second);

    hdf5::Group &runtime = file.root().create_group("RuntimePars");
    {
      char stamp[64];
      const time_t now = ::time(nullptr);
      struct tm tmv;
      localtime_r(&now, &tmv);
      strftime(stamp, sizeof(stamp), "%d/%m/%Y, %H:%M:%S", &tmv); /* Utilities::get_timestamp */
      runtime.write_attribute("Creation time", stamp);
    }
    runtime.write_attribute("Iteration", uint32_t(iteration));

    hdf5::Group &units = file.root().create_group("Units");
    units.write_attribute("Unit current in cgs (U_I)", 1.);
    units.write_attribute("Unit length in cgs (U_L)", 100.);
    units.write_attribute("Unit mass in cgs (U_M)", 1000.);
    units.write_attribute("Unit temperature in cgs (U_T)", 1.);
    units.write_attribute("Unit time in cgs (U_t)", 1.);

    hdf5::Group &part = file.root().create_group("PartType0");
    std::vector<double> coordinates;
    if (fields_.coordinates) {
      coordinates.resize(3 * n);
      const Vec3 &anchor = grid.get_box_anchor();
      for (size_t i = 0; i < n; ++i) {
        const Vec3 x = grid.get_cell_midpoint(i);
        for (int k = 0; k < 3; ++k) coordinates[3 * i + k] = x[k] - anchor[k];
      }
      part.create_dataset("Coordinates", hdf5::Type::F64, {n, 3}, coordinates.data());
    }
    if (fields_.number_density) part.create_dataset("NumberDensity", hdf5::Type::F64, {n}, grid.number_density.data());
    if (fields_.temperature) part.create_dataset("Temperature", hdf5::Type::F64, {n}, grid.temperature.data());
    for (int ion = 0; ion < CMIB_NUM_IONS; ++ion)
      if (fields_.ion_present(ion))
        part.create_dataset(std::string("NeutralFraction") + ion_symbol(ion), hdf5::Type::F64, {n},
                            grid.ionic_fraction.data() + (size_t)ion * n);
    file.write(filename(iteration));
  }

private:
  std::string prefix_, folder_;
  DensityGridWriterFields fields_;
  uint32_t padding_;
};

struct DensityGridWriterFactory {
  /* DensityGridWriterFactory.hpp:86-110; the default type is Gadget, as in the reference */
  static DensityGridWriter *generate(const std::string &output_folder, ParameterFile &params, Log *log = nullptr) {
    const std::string type = params.get_value<std::string>("DensityGridWriter:type", "Gadget");
    if (log) log->write_info("Requested DensityGridWriter type: ", type);
    if (type == "AsciiFile") return new AsciiFileDensityGridWriter(output_folder, params);
    if (type == "Gadget") return new GadgetDensityGridWriter(output_folder, params);
    cmi_error("Unknown DensityGridWriter type: \"%s\".", type.c_str());
  }
};

/* ---- NCCL, bound at run time ----
 * Only multi-GPU runs need NCCL, and a process may already carry one (PyTorch bundles its own
 * libnccl.so.2): dlopen picks up whatever is loaded, else the system library; nothing is linked. */
struct NcclApi {
  decltype(&ncclCommInitAll) CommInitAll = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  static NcclApi &get() {
    static NcclApi api = load();
    return api;
  }

private:
  static NcclApi load() {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) cmi_error("Multi-GPU runs need NCCL: %s", dlerror());
    NcclApi a;
    a.CommInitAll = reinterpret_cast<decltype(a.CommInitAll)>(dlsym(h, "ncclCommInitAll"));
    a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(dlsym(h, "ncclAllReduce"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    if (!a.CommInitAll || !a.AllReduce || !a.CommDestroy) cmi_error("libnccl lacks the expected entry points!");
    return a;
  }
};

/* ---- the driver ---- */
class IonizationSimulation {
public:
  /* same leading arguments as the reference (IonizationSimulation.hpp:196-202); num_thread is
   * accepted and ignored (the parallelism is the GPU's); the MPICommunicator* is replaced by the
   * list of devices of this node.  With more than one device the packets of an iteration are
   * split by global packet id, every device holds the whole grid, and ONE ncclAllReduce over the
   * accumulator buffers replaces the reference's 16 chunked MPI_Allreduce + counter reductions
   * (IonizationSimulation.cpp:410-416, 458-529); the state update is then run on every device
   * (replicated), so no gather is needed (:540-618). */
  /* task_based = true: the parameter surface of the reference's other driver (`CMacIonize --task-based`,
   * TaskBasedIonizationSimulation.cpp:190-370) on the same GPU path: Monte Carlo parameters come from the
   * `TaskBasedIonizationSimulation:` block (number of iterations 10, number of photons 1e6, random seed,
   * output folder, diffuse field), the periodicity from `DensitySubGridCreator:periodicity`, and the
   * diffuse re-emission handler exists only when `diffuse field` is true (:338-346).  The task queues,
   * buffers, subgrids and source copies of that driver are CPU scheduling and have no counterpart here
   * (their keys are read so that they show up in the used-values file).  The packet conventions that
   * differ inside the reference's task-based code (abundance-weighted cross sections carried by the
   * packet, DensitySubGrid.hpp:593-612) describe the same physics; results agree with either reference
   * driver within Monte Carlo noise (tests/test_gpu_benchmarks.py). */
  IonizationSimulation(bool write_output, bool every_iteration_output, bool output_statistics, int num_thread,
                       const std::string &parameterfile, const std::vector<int> &devices, Log *log = nullptr,
                       bool task_based = false)
      : every_iteration_output_(every_iteration_output), output_statistics_(output_statistics), log_(log),
        parameter_file_(parameterfile), block_(task_based ? "TaskBasedIonizationSimulation:" : "IonizationSimulation:"),
        number_of_iterations_(parameter_file_.get_value<uint32_t>(block_ + "number of iterations", 10)),
        number_of_photons_(parameter_file_.get_value<uint64_t>(block_ + "number of photons", task_based ? 1000000 : 100000)),
        number_of_photons_init_(task_based ? number_of_photons_
                                           : parameter_file_.get_value<uint64_t>(block_ + "number of photons first loop",
                                                                                 number_of_photons_)),
        abundances_(Abundances::generate(parameter_file_, log)), devices_(devices) {
    (void)num_thread;
    if (devices_.empty()) cmi_error("No device given!");
    if (task_based) {
      (void)parameter_file_.get_value<uint32_t>(block_ + "source copy level", 4);
      (void)parameter_file_.get_value<uint64_t>(block_ + "number of buffers", 50000);
      (void)parameter_file_.get_value<uint64_t>(block_ + "queue size per thread", 10000);
      (void)parameter_file_.get_value<uint64_t>(block_ + "shared queue size", 100000);
      (void)parameter_file_.get_value<uint64_t>(block_ + "number of tasks", 500000);
    }
    cross_sections_.reset(CrossSections::generate(parameter_file_, log_));
    recombination_rates_.reset(RecombinationRates::generate(parameter_file_, log_));
    density_function_.reset(DensityFunctionFactory::generate(parameter_file_, log_));
    density_mask_.reset(DensityMaskFactory::generate(parameter_file_, log_));
    SimulationBox box(parameter_file_);
    if (task_based) { /* DensitySubGridCreator(box, params) (DensitySubGridCreator.hpp:106-117) */
      (void)parameter_file_.get_value<std::array<int32_t, 3>>("DensitySubGridCreator:number of subgrids", {8, 8, 8});
      box.periodicity = parameter_file_.get_value<std::array<bool, 3>>("DensitySubGridCreator:periodicity", {false, false, false});
    }
    const std::string grid_type = parameter_file_.get_value<std::string>("DensityGrid:type", "Cartesian");
    if (grid_type != "Cartesian")
      cmi_error("Unknown DensityGrid type: \"%s\" (the B200 backend provides Cartesian)!", grid_type.c_str());
    const auto ncell = parameter_file_.get_value<std::array<int32_t, 3>>("DensityGrid:number of cells", {64, 64, 64});
    for (int device : devices_) density_grids_.emplace_back(new CartesianDensityGrid(box, ncell, device));
    photon_source_distribution_.reset(PhotonSourceDistributionFactory::generate(parameter_file_, log_));
    photon_source_spectrum_.reset(PhotonSourceSpectrum::generate("PhotonSourceSpectrum", parameter_file_, log_));
    if (photon_source_distribution_ && !photon_source_spectrum_)
      cmi_error("No spectrum provided for the discrete photon sources!");
    /* ContinuousPhotonSourceFactory (src/ContinuousPhotonSourceFactory.hpp:69-100) and its spectrum
     * (IonizationSimulation.cpp:164-174) */
    const std::string continuous_type = parameter_file_.get_value<std::string>("ContinuousPhotonSource:type", "None");
    if (log_) log_->write_info("Requested ContinuousPhotonSource type: ", continuous_type, ".");
    if (continuous_type != "None" && continuous_type != "Isotropic" && continuous_type != "Planar" &&
        continuous_type != "DistantStar" && continuous_type != "ExtendedDisc" && continuous_type != "SpiralGalaxy")
      cmi_error("Unknown ContinuousPhotonSource type: \"%s\" (the B200 backend provides Isotropic, Planar, DistantStar, "
                "ExtendedDisc and SpiralGalaxy)!",
                continuous_type.c_str());
    /* DistantStarContinuousPhotonSource(box, params) (src/DistantStarContinuousPhotonSource.hpp:92-97) */
    Vec3 star_position = {0., 0., 0.};
    double star_area = 0.;
    if (continuous_type == "DistantStar") {
      star_position = parameter_file_.get_physical_vector<QUANTITY_LENGTH>("ContinuousPhotonSource:position");
      /* get_total_surface_area (:194-212): the faces turned towards the star */
      const bool ex[3] = {star_position[0] < box.anchor[0] || star_position[0] > box.anchor[0] + box.sides[0],
                          star_position[1] < box.anchor[1] || star_position[1] > box.anchor[1] + box.sides[1],
                          star_position[2] < box.anchor[2] || star_position[2] > box.anchor[2] + box.sides[2]};
      if (ex[0]) star_area += box.sides[1] * box.sides[2];
      if (ex[1]) star_area += box.sides[0] * box.sides[2];
      if (ex[2]) star_area += box.sides[0] * box.sides[1];
      if (!ex[0] && !ex[1] && !ex[2]) cmi_error("External stellar source lies inside the simulation box. This will not work!");
    }
    /* PlanarContinuousPhotonSource(ParameterFile&) (src/PlanarContinuousPhotonSource.hpp:133-151) */
    int planar_axis = 2;
    double planar_intercept = 0., planar_anchor[2] = {0., 0.}, planar_sides[2] = {1., 1.}, planar_luminosity = 0.;
    if (continuous_type == "Planar") {
      const std::string axis = parameter_file_.get_value<std::string>("ContinuousPhotonSource:normal axis", "z");
      if (axis == "x") planar_axis = 0;
      else if (axis == "y") planar_axis = 1;
      else if (axis == "z") planar_axis = 2;
      else cmi_error("Unknown coordinate axis name: \"%s\"!", axis.c_str());
      planar_intercept = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:intercept", "0. m");
      planar_anchor[0] = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:anchor 0", "0. m");
      planar_anchor[1] = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:anchor 1", "0. m");
      planar_sides[0] = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:side 0", "1. m");
      planar_sides[1] = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:side 1", "1. m");
      planar_luminosity = parameter_file_.get_physical_value<QUANTITY_FREQUENCY>("ContinuousPhotonSource:luminosity", "1.e48 s^-1");
    }
    /* ExtendedDiscContinuousPhotonSource(box, params) (src/ExtendedDiscContinuousPhotonSource.hpp:102-117) */
    double disc_scale_height = 0.;
    if (continuous_type == "ExtendedDisc") {
      const std::string axis = parameter_file_.get_value<std::string>("ContinuousPhotonSource:normal axis", "z");
      if (axis == "x") planar_axis = 0;
      else if (axis == "y") planar_axis = 1;
      else if (axis == "z") planar_axis = 2;
      else cmi_error("Unknown coordinate axis name: \"%s\"!", axis.c_str());
      planar_intercept = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:intercept", "0. m");
      disc_scale_height = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:scale height", "200. pc");
      planar_luminosity = parameter_file_.get_physical_value<QUANTITY_FREQUENCY>("ContinuousPhotonSource:luminosity", "1.e48 s^-1");
    }
    /* SpiralGalaxyContinuousPhotonSource(box, params) (src/SpiralGalaxyContinuousPhotonSource.hpp:107-120) */
    double galaxy_r_stars = 0., galaxy_h_stars = 0., galaxy_B_over_T = 0.;
    if (continuous_type == "SpiralGalaxy") {
      galaxy_r_stars = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:scale length stars", "5. kpc");
      galaxy_h_stars = parameter_file_.get_physical_value<QUANTITY_LENGTH>("ContinuousPhotonSource:scale height stars", "0.6 kpc");
      galaxy_B_over_T = parameter_file_.get_value<double>("ContinuousPhotonSource:bulge over total ratio", 0.2);
    }
    continuous_photon_source_spectrum_.reset(
        PhotonSourceSpectrum::generate("ContinuousPhotonSourceSpectrum", parameter_file_, log_));
    const bool has_continuous = (continuous_type != "None");
    if (has_continuous && !continuous_photon_source_spectrum_)
      cmi_error("No spectrum provided for the continuous photon sources!");
    if (!photon_source_distribution_ && !has_continuous) cmi_error("No photon sources!");
    double continuous_luminosity = 0.;
    if (continuous_type == "Planar" || continuous_type == "ExtendedDisc") {
      continuous_luminosity = planar_luminosity; /* has_total_luminosity() (PhotonSource.cpp:101-103) */
    } else if (has_continuous) {
      /* PhotonSource.cpp:104-108: total surface area (IsotropicContinuousPhotonSource.hpp:187-192) x total flux */
      if (continuous_photon_source_spectrum_->total_flux < 0.) cmi_error("This function should not be used!");
      const double area = (continuous_type == "DistantStar")
                              ? star_area
                          : (continuous_type == "SpiralGalaxy")
                              ? 1. /* SpiralGalaxyContinuousPhotonSource::get_total_surface_area (:194) */
                              : 2. * box.sides[0] * box.sides[1] + 2. * box.sides[0] * box.sides[2] +
                                    2. * box.sides[1] * box.sides[2];
      continuous_luminosity = area * continuous_photon_source_spectrum_->total_flux;
    }
    /* classic driver: always through the factory; task-based driver: only when the diffuse field is switched
     * on (TaskBasedIonizationSimulation.cpp:338-346; the second condition deals with old parameter files) */
    if (!task_based || parameter_file_.get_value<bool>(block_ + "diffuse field", false) ||
        parameter_file_.has_value("PhotonSource:diffuse field"))
      reemission_ = DiffuseReemissionHandler::generate(parameter_file_, log_);
    const cmib_temperature_params tp = temperature_calculator_parameters(parameter_file_);

    /* configure every device context alike: PhotonSource ctor (PhotonSource.cpp:55-146) */
    const size_t ns = photon_source_distribution_ ? photon_source_distribution_->get_number_of_sources() : 0;
    std::vector<double> pos(3 * ns), w(ns);
    for (size_t i = 0; i < ns; ++i) {
      const Vec3 p = photon_source_distribution_->get_position(i);
      pos[3 * i] = p[0]; pos[3 * i + 1] = p[1]; pos[3 * i + 2] = p[2];
      w[i] = photon_source_distribution_->get_weight(i);
    }
    const double discrete_luminosity = photon_source_distribution_ ? photon_source_distribution_->get_total_luminosity() : 0.;
    total_luminosity_ = discrete_luminosity + continuous_luminosity;
    for (auto &grid : density_grids_) {
      cmib_context *ctx = grid->context();
      CMIB_CALL(cmib_set_abundances(ctx, abundances_.abundance));
      CMIB_CALL(cross_sections_->set_on(ctx));
      CMIB_CALL(cmib_set_recombination_rates(ctx, recombination_rates_->kind, recombination_rates_->fixed));
      CMIB_CALL(cmib_set_sources(ctx, (int32_t)ns, pos.data(), w.data(), discrete_luminosity));
      if (ns > 0) CMIB_CALL(photon_source_spectrum_->set_on(ctx, 0));
      if (has_continuous) {
        CMIB_CALL(continuous_photon_source_spectrum_->set_on(ctx, 1));
        if (continuous_type == "Planar")
          CMIB_CALL(cmib_set_planar_source_geometry(ctx, planar_axis, planar_intercept, planar_anchor, planar_sides));
        if (continuous_type == "DistantStar") CMIB_CALL(cmib_set_distant_star_position(ctx, star_position.data()));
        if (continuous_type == "ExtendedDisc")
          CMIB_CALL(cmib_set_extended_disc_geometry(ctx, planar_axis, planar_intercept, disc_scale_height));
        if (continuous_type == "SpiralGalaxy")
          CMIB_CALL(cmib_set_spiral_galaxy_geometry(ctx, galaxy_r_stars, galaxy_h_stars, galaxy_B_over_T));
        CMIB_CALL(cmib_set_continuous_source(ctx,
                                             continuous_type == "Planar" ? CMIB_CONTINUOUS_PLANAR
                                             : continuous_type == "DistantStar" ? CMIB_CONTINUOUS_DISTANT_STAR
                                             : continuous_type == "ExtendedDisc" ? CMIB_CONTINUOUS_EXTENDED_DISC
                                             : continuous_type == "SpiralGalaxy" ? CMIB_CONTINUOUS_SPIRAL_GALAXY
                                                                                : CMIB_CONTINUOUS_ISOTROPIC,
                                             continuous_luminosity,
                                             continuous_photon_source_spectrum_->kind,
                                             continuous_photon_source_spectrum_->param));
      }
      CMIB_CALL(cmib_set_reemission(ctx, reemission_.kind, reemission_.probability, reemission_.frequency));
      CMIB_CALL(cmib_set_temperature_params(ctx, &tp));
    }
    if (devices_.size() > 1) {
      comms_.resize(devices_.size());
      if (NcclApi::get().CommInitAll(comms_.data(), (int)devices_.size(), devices_.data()) != ncclSuccess)
        cmi_error("ncclCommInitAll failed for %zu devices!", devices_.size());
    }

    output_folder_ = parameter_file_.get_value<std::string>(block_ + "output folder", ".");
    if (write_output) {
      density_grid_writer_.reset(DensityGridWriterFactory::generate(output_folder_, parameter_file_, log_));
    }
    random_seed_ = parameter_file_.get_value<int32_t>(block_ + "random seed", 42);
    if (parameter_file_.get_value<bool>(block_ + "enable trackers", false))
      cmi_error("Trackers are not provided by the B200 backend!");
    if (write_output) {
      std::ofstream pfile(parameterfile + ".used-values");
      parameter_file_.print_contents(pfile);
      if (log_) log_->write_status("Wrote used parameters to ", parameterfile + ".used-values", ".");
    }
  }

  IonizationSimulation(bool write_output, bool every_iteration_output, bool output_statistics, int num_thread,
                       const std::string &parameterfile, int device = 0, Log *log = nullptr, bool task_based = false)
      : IonizationSimulation(write_output, every_iteration_output, output_statistics, num_thread, parameterfile,
                             std::vector<int>{device}, log, task_based) {}

  ~IonizationSimulation() {
    for (ncclComm_t c : comms_) NcclApi::get().CommDestroy(c);
  }

  /* IonizationSimulation::initialize (IonizationSimulation.cpp:239-326) */
  void initialize(DensityFunction *density_function = nullptr) {
    if (!density_function) density_function = density_function_.get();
    density_function->initialize();
    density_grids_[0]->initialize(*density_function);
    if (density_mask_) { /* IonizationSimulation.cpp:308-321 */
      if (log_) log_->write_status("Initializing DensityMask...");
      density_mask_->initialize();
      if (log_) log_->write_status("Done initializing mask. Applying mask...");
      density_mask_->apply(*density_grids_[0]);
      density_grids_[0]->upload();
      if (log_) log_->write_status("Done applying mask.");
    }
    for (size_t d = 1; d < density_grids_.size(); ++d) {
      density_grids_[d]->number_density = density_grids_[0]->number_density;
      density_grids_[d]->temperature = density_grids_[0]->temperature;
      density_grids_[d]->ionic_fraction = density_grids_[0]->ionic_fraction;
      density_grids_[d]->upload();
    }
  }

  struct IterationResult {
    double totweight = 0.;
    double typecount[CMIB_NUM_PACKET_TYPES] = {0., 0., 0., 0.};
    double shoot_seconds = 0., update_seconds = 0.;
  };

  /* one pass of the loop body of IonizationSimulation::run (IonizationSimulation.cpp:359-643) */
  IterationResult iteration(uint32_t loop, uint64_t numphoton) {
    using clock = std::chrono::steady_clock;
    const size_t ndev = density_grids_.size();
    std::vector<IterationResult> part(ndev);
    std::vector<std::string> errors(ndev);
    auto work = [&](size_t d) {
      try {
        cmib_context *ctx = density_grids_[d]->context();
        /* MPICommunicator::distribute (MPICommunicator.hpp:207-222): contiguous id blocks */
        const uint64_t per = numphoton / ndev;
        const uint64_t lo = d * per;
        const uint64_t cnt = (d + 1 < ndev) ? per : numphoton - lo;
        IterationResult &r = part[d];
        density_grids_[d]->reset_grid();
        CMIB_CALL(cmib_update_reemission_probabilities(ctx));
        CMIB_CALL(cmib_synchronize(ctx));
        const auto t0 = clock::now();
        CMIB_CALL(cmib_shoot(ctx, cnt, lo, (uint64_t)(int64_t)random_seed_, loop, &r.totweight, r.typecount));
        const auto t1 = clock::now();
        double totweight = r.totweight;
        if (ndev > 1) {
          void *buf = nullptr, *stream = nullptr;
          uint64_t n = 0;
          CMIB_CALL(cmib_accumulator_buffer(ctx, &buf, &n));
          CMIB_CALL(cmib_stream(ctx, &stream));
          if (NcclApi::get().AllReduce(buf, buf, n, ncclDouble, ncclSum, comms_[d], (cudaStream_t)stream) != ncclSuccess)
            cmi_error("ncclAllReduce failed on device %d!", devices_[d]);
          totweight = 0.; /* use the reduced device-side sum */
        }
        CMIB_CALL(cmib_update_state(ctx, loop, totweight));
        CMIB_CALL(cmib_synchronize(ctx));
        const auto t2 = clock::now();
        r.shoot_seconds = std::chrono::duration<double>(t1 - t0).count();
        r.update_seconds = std::chrono::duration<double>(t2 - t1).count();
      } catch (const std::exception &e) {
        errors[d] = e.what();
      }
    };
    if (ndev == 1) {
      work(0);
    } else {
      std::vector<std::thread> threads;
      for (size_t d = 0; d < ndev; ++d) threads.emplace_back(work, d);
      for (auto &t : threads) t.join();
    }
    for (const std::string &e : errors)
      if (!e.empty()) throw Error(e);
    IterationResult r;
    for (const IterationResult &p : part) {
      r.totweight += p.totweight;
      for (int t = 0; t < CMIB_NUM_PACKET_TYPES; ++t) r.typecount[t] += p.typecount[t];
      r.shoot_seconds = std::max(r.shoot_seconds, p.shoot_seconds);
      r.update_seconds = std::max(r.update_seconds, p.update_seconds);
    }
    return r;
  }

  /* IonizationSimulation::run */
  /* external_writer: IonizationSimulation::run(DensityGridWriter*) (IonizationSimulation.cpp:334, 655-659):
   * called once with the final grid (host mirror refreshed) */
  void run(const std::function<void(CartesianDensityGrid &)> &external_writer = nullptr) {
    CartesianDensityGrid &grid = *density_grids_[0];
    if (density_grid_writer_) { grid.download(); density_grid_writer_->write(grid, 0, parameter_file_); }
    double shoot = 0., update = 0.;
    for (uint32_t loop = 0; loop < number_of_iterations_; ++loop) {
      if (log_) log_->write_status("Starting loop ", loop, ".");
      const uint64_t lnumphoton = (loop == 0) ? number_of_photons_init_ : number_of_photons_;
      if (log_) log_->write_status("Start shooting ", lnumphoton, " photons...");
      const IterationResult r = iteration(loop, lnumphoton);
      shoot += r.shoot_seconds;
      update += r.update_seconds;
      if (log_) log_->write_status("Done shooting photons.");
      if (output_statistics_ && log_) {
        /* IonizationSimulation.cpp:421-446 */
        const double W = r.totweight;
        log_->write_info(100. * r.typecount[3] / W, "% of photons were reemitted as non-ionizing photons.");
        log_->write_info(100. * (r.typecount[1] + r.typecount[2]) / W, "% of photons were scattered.");
        const double escape = 100. * (W - r.typecount[3]) / W;
        log_->write_info("Escape fraction: ", escape, "%.");
        log_->write_info("Escape fraction from diffuse hydrogen: ", 100. * r.typecount[1] / W, "%.");
        log_->write_info("Escape fraction from diffuse helium: ", 100. * r.typecount[2] / W, "%.");
      }
      if (every_iteration_output_ && density_grid_writer_ && loop + 1 < number_of_iterations_)
      { grid.download(); density_grid_writer_->write(grid, loop + 1, parameter_file_); }
    }
    if (density_grid_writer_) { grid.download(); density_grid_writer_->write(grid, number_of_iterations_, parameter_file_); }
    if (external_writer) {
      grid.download();
      external_writer(grid);
    }
    if (log_) {
      log_->write_status("Total photon shooting time: ", shoot, " s.");
      log_->write_status("Total cell update time: ", update, " s.");
    }
    total_shoot_seconds_ = shoot;
    total_update_seconds_ = update;
  }

  CartesianDensityGrid &get_density_grid(size_t device_index = 0) { return *density_grids_[device_index]; }
  size_t get_number_of_devices() const { return density_grids_.size(); }
  ParameterFile &get_parameter_file() { return parameter_file_; }
  uint32_t get_number_of_iterations() const { return number_of_iterations_; }
  uint64_t get_number_of_photons() const { return number_of_photons_; }
  double get_total_luminosity() const { return total_luminosity_; }
  double total_shoot_seconds() const { return total_shoot_seconds_; }
  double total_update_seconds() const { return total_update_seconds_; }

private:
  bool every_iteration_output_, output_statistics_;
  Log *log_;
  ParameterFile parameter_file_;
  std::string block_; /* "IonizationSimulation:" or "TaskBasedIonizationSimulation:" */
  uint32_t number_of_iterations_;
  uint64_t number_of_photons_, number_of_photons_init_;
  Abundances abundances_;
  std::vector<int> devices_;
  std::unique_ptr<CrossSections> cross_sections_;
  std::unique_ptr<RecombinationRates> recombination_rates_;
  std::unique_ptr<DensityFunction> density_function_;
  std::vector<std::unique_ptr<CartesianDensityGrid>> density_grids_;
  std::vector<ncclComm_t> comms_;
  std::unique_ptr<PhotonSourceDistribution> photon_source_distribution_;
  std::unique_ptr<PhotonSourceSpectrum> photon_source_spectrum_;
  std::unique_ptr<PhotonSourceSpectrum> continuous_photon_source_spectrum_;
  std::unique_ptr<FractalDensityMask> density_mask_;
  DiffuseReemissionHandler reemission_;
  std::unique_ptr<DensityGridWriter> density_grid_writer_;
  std::string output_folder_;
  double total_luminosity_ = 0.;
  int32_t random_seed_ = 42;
  double total_shoot_seconds_ = 0., total_update_seconds_ = 0.;
};

} // namespace cmi
