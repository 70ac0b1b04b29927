/*
 * PhotonSourceDistributions.hpp — PhotonSourceDistribution plugins of the host layer (PhotonSourceDistributionFactory.hpp): positions, weights
 * and total luminosity of the discrete sources, handed to cmib_set_sources.
 * Part of the host layer described in IonizationSimulation.hpp (class map, reference citations).
 */
#pragma once
#include "HostCommon.hpp"

namespace cmi {

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

} // namespace cmi
