/*
 * DevicePlugins.hpp — Plugins that are pure parameters of the device path: PhotonSourceSpectrum (+ masks, tabulated spectra),
 * CrossSections, RecombinationRates, AbundanceModel, DiffuseReemissionHandler, TemperatureCalculator parameters.
 * Part of the host layer described in IonizationSimulation.hpp (class map, reference citations).
 */
#pragma once
#include "HostCommon.hpp"

namespace cmi {

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

} // namespace cmi
