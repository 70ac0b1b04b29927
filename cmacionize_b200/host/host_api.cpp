/*
 * host_api.cpp — C entry points over the C++ host layer (IonizationSimulation.hpp), so that
 * the tests and bench.py drive the SAME host code a C++ caller links against.
 *
 * Mirrors the shape of the reference's coarse library interface (cmi_init / cmi_destroy,
 * /root/reference/src/CMILibrary.cpp:48-58) but per-object instead of global singletons.
 * Errors: return code + cmih_last_error(); with CMIB_ABORT_ON_ERROR=1 the message is printed
 * and the process aborts like the reference's cmac_error.
 */
#include <cstdlib>
#include <cstring>
#include <string>

#include "IonizationSimulation.hpp"
#include "SPHArrayInterface.hpp"

using namespace cmi;

namespace {
thread_local std::string g_error;
int fail(const std::exception &e) {
  g_error = e.what();
  const char *a = getenv("CMIB_ABORT_ON_ERROR");
  if (a && a[0] == '1') {
    fprintf(stderr, "%s\n", g_error.c_str());
    abort();
  }
  return 1;
}
struct Sim {
  Log log;
  IonizationSimulation sim;
  static std::vector<int> device_list(int first, int n) {
    std::vector<int> v;
    for (int g = 0; g < (n > 0 ? n : 1); ++g) v.push_back(first + g);
    return v;
  }
  Sim(const char *paramfile, int device, int ngpus, int write_output, int verbose, bool task_based = false)
      : log(verbose ? Log::INFO : Log::WARNING),
        sim(write_output != 0, false, verbose != 0, -1, paramfile, device_list(device, ngpus), &log, task_based) {}
};
} // namespace

#define CMIH_TRY(body)                      \
  try {                                     \
    body;                                   \
    return 0;                               \
  } catch (const std::exception &e) {       \
    return fail(e);                         \
  }

extern "C" {

const char *cmih_last_error(void) { return g_error.c_str(); }

/* IonizationSimulation(write_output, false, output_statistics, -1, parameterfile, device, log) */
int cmih_simulation_create(const char *paramfile, int device, int write_output, int verbose, void **out) {
  CMIH_TRY(*out = new Sim(paramfile, device, 1, write_output, verbose));
}
/* devices device .. device+ngpus-1 of this node; one NCCL all-reduce per iteration */
int cmih_simulation_create_multi(const char *paramfile, int device, int ngpus, int write_output, int verbose,
                                 void **out) {
  CMIH_TRY(*out = new Sim(paramfile, device, ngpus, write_output, verbose));
}
/* the `CMacIonize --task-based` parameter surface (TaskBasedIonizationSimulation: block) on the same path */
int cmih_simulation_create_task_based(const char *paramfile, int device, int ngpus, int write_output, int verbose,
                                      void **out) {
  CMIH_TRY(*out = new Sim(paramfile, device, ngpus, write_output, verbose, true));
}
int cmih_simulation_destroy(void *h) {
  delete static_cast<Sim *>(h);
  return 0;
}
int cmih_simulation_initialize(void *h) { CMIH_TRY(static_cast<Sim *>(h)->sim.initialize()); }
int cmih_simulation_run(void *h) { CMIH_TRY(static_cast<Sim *>(h)->sim.run()); }
/* out[8]: totweight, typecount[4], shoot seconds, update seconds, 0 */
int cmih_simulation_iteration(void *h, uint32_t loop, uint64_t numphoton, double *out) {
  CMIH_TRY({
    const IonizationSimulation::IterationResult r = static_cast<Sim *>(h)->sim.iteration(loop, numphoton);
    out[0] = r.totweight;
    for (int t = 0; t < 4; ++t) out[1 + t] = r.typecount[t];
    out[5] = r.shoot_seconds;
    out[6] = r.update_seconds;
    out[7] = 0.;
  });
}
/* info[4]: number of cells, iterations, photons per iteration, total luminosity */
int cmih_simulation_info(void *h, double *info) {
  CMIH_TRY({
    IonizationSimulation &s = static_cast<Sim *>(h)->sim;
    info[0] = (double)s.get_density_grid().get_number_of_cells();
    info[1] = (double)s.get_number_of_iterations();
    info[2] = (double)s.get_number_of_photons();
    info[3] = s.get_total_luminosity();
  });
}
/* the cmib_context of the simulation's grid (for cmib_download_cells etc.) */
int cmih_simulation_context(void *h, cmib_context **ctx) {
  CMIH_TRY(*ctx = static_cast<Sim *>(h)->sim.get_density_grid().context());
}
int cmih_simulation_context_of(void *h, int device_index, cmib_context **ctx) {
  CMIH_TRY(*ctx = static_cast<Sim *>(h)->sim.get_density_grid((size_t)device_index).context());
}
/* multi-GPU: bring the per-cell state that stays with the owners of the cell chunks (metal fractions, heating terms) to
 * every device, before the cells of any device are downloaded */
int cmih_simulation_gather_state(void *h) { CMIH_TRY(static_cast<Sim *>(h)->sim.gather_state()); }

/* ---- parameter-file probes (parity tests against the reference's ParameterFile) ---- */
int cmih_paramfile_open(const char *filename, void **out) { CMIH_TRY(*out = new ParameterFile(filename)); }
int cmih_paramfile_close(void *h) {
  delete static_cast<ParameterFile *>(h);
  return 0;
}
/* quantity: the Quantity enum of ParameterFile.hpp; unit conversion to SI */
int cmih_convert_to_SI(int quantity, double value, const char *unit, double *out) {
  CMIH_TRY(*out = UnitConverter::to_SI((Quantity)quantity, value, unit));
}
int cmih_convert(double value, const char *unit_from, const char *unit_to, double *out) {
  CMIH_TRY(*out = UnitConverter::convert(value, unit_from, unit_to));
}
int cmih_paramfile_get_string(void *h, const char *key, const char *default_value, char *out, int n) {
  CMIH_TRY({
    const std::string v = static_cast<ParameterFile *>(h)->get_value(std::string(key), default_value);
    strncpy(out, v.c_str(), n - 1);
    out[n - 1] = 0;
  });
}
int cmih_paramfile_get_double(void *h, const char *key, double default_value, double *out) {
  CMIH_TRY(*out = static_cast<ParameterFile *>(h)->get_value<double>(key, default_value));
}
int cmih_paramfile_get_physical(void *h, int quantity, const char *key, const char *default_value, double *out) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    switch ((Quantity)quantity) {
    case QUANTITY_LENGTH: *out = p.get_physical_value<QUANTITY_LENGTH>(key, default_value); break;
    case QUANTITY_NUMBER_DENSITY: *out = p.get_physical_value<QUANTITY_NUMBER_DENSITY>(key, default_value); break;
    case QUANTITY_TEMPERATURE: *out = p.get_physical_value<QUANTITY_TEMPERATURE>(key, default_value); break;
    case QUANTITY_FREQUENCY: *out = p.get_physical_value<QUANTITY_FREQUENCY>(key, default_value); break;
    case QUANTITY_SURFACE_AREA: *out = p.get_physical_value<QUANTITY_SURFACE_AREA>(key, default_value); break;
    case QUANTITY_REACTION_RATE: *out = p.get_physical_value<QUANTITY_REACTION_RATE>(key, default_value); break;
    default: cmi_error("quantity %d not exposed by this probe", quantity);
    }
  });
}
/* the used-values dump (YAMLDictionary::print_contents(stream, true)) into a caller buffer */
int cmih_paramfile_used_values(void *h, char *out, int n) {
  CMIH_TRY({
    std::ostringstream s;
    static_cast<YAMLDictionary *>(static_cast<ParameterFile *>(h))->print_contents(s, true);
    strncpy(out, s.str().c_str(), n - 1);
    out[n - 1] = 0;
  });
}
/* evaluate the parameter file's DensityFunction at n points (x[n][3]) -> number density,
 * temperature, neutral fraction of H */
int cmih_density_function(void *h, int64_t n, const double *x, double *dens, double *temp, double *xH) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    std::unique_ptr<DensityFunction> f(DensityFunctionFactory::generate(p));
    f->initialize();
    for (int64_t i = 0; i < n; ++i) {
      const DensityValues v = (*f)({x[3 * i], x[3 * i + 1], x[3 * i + 2]});
      dens[i] = v.number_density;
      temp[i] = v.temperature;
      xH[i] = v.ionic_fraction[0];
    }
  });
}

/* the parameter file's PhotonSourceSpectrum for `role` ("PhotonSourceSpectrum" /
 * "ContinuousPhotonSourceSpectrum"): info = {kind, param, total flux, table length}; the table of a
 * tabulated spectrum is copied to freq / cdf (capacity entries each) */
int cmih_photon_source_spectrum(void *h, const char *role, double *info, double *freq, double *cdf, int capacity) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    std::unique_ptr<PhotonSourceSpectrum> s(PhotonSourceSpectrum::generate(role, p));
    if (!s) throw std::runtime_error("no spectrum (type None)");
    info[0] = s->kind;
    info[1] = s->param;
    info[2] = s->total_flux;
    info[3] = (double)s->frequencies.size();
    if ((int)s->frequencies.size() > capacity) throw std::runtime_error("table capacity too small");
    for (size_t i = 0; i < s->frequencies.size(); ++i) {
      freq[i] = s->frequencies[i];
      cdf[i] = s->cumulative_distribution[i];
    }
  });
}

/* the host stage of IonizationSimulation::initialize without a device: DensityFunction evaluated on
 * the parameter file's Cartesian grid, then the DensityMask (if any) -> number density per cell */
int cmih_initial_number_density(void *h, int64_t n, double *dens) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    std::unique_ptr<DensityFunction> f(DensityFunctionFactory::generate(p));
    std::unique_ptr<FractalDensityMask> mask(DensityMaskFactory::generate(p));
    const SimulationBox box(p);
    CartesianCells cells(box, p.get_value<std::array<int32_t, 3>>("DensityGrid:number of cells", {64, 64, 64}));
    if ((int64_t)cells.get_number_of_cells() != n) throw std::runtime_error("wrong number of cells");
    f->initialize();
    cells.set_densities(*f);
    if (mask) {
      mask->initialize();
      mask->apply(cells);
    }
    for (int64_t i = 0; i < n; ++i) dens[i] = cells.number_density[i];
  });
}
/* the parameter file's DensityGridWriter on given cell arrays (no device): number density [n], temperature [n],
 * ionic fractions [14][n] on the parameter file's Cartesian grid -> writes the snapshot, returns its file name */
int cmih_write_snapshot(void *h, const char *output_folder, uint32_t iteration, double time, int64_t n, const double *dens,
                        const double *temp, const double *fractions, char *filename, int nfilename) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    const SimulationBox box(p);
    p.get_value<std::string>("DensityGrid:type", "Cartesian"); /* as the driver does: part of the used values */
    CartesianCells cells(box, p.get_value<std::array<int32_t, 3>>("DensityGrid:number of cells", {64, 64, 64}));
    if ((int64_t)cells.get_number_of_cells() != n) throw std::runtime_error("wrong number of cells");
    std::copy(dens, dens + n, cells.number_density.begin());
    std::copy(temp, temp + n, cells.temperature.begin());
    std::copy(fractions, fractions + (size_t)CMIB_NUM_IONS * n, cells.ionic_fraction.begin());
    std::unique_ptr<DensityGridWriter> writer(DensityGridWriterFactory::generate(output_folder, p));
    writer->write(cells, iteration, p, time);
    std::string name;
    if (auto *g = dynamic_cast<GadgetDensityGridWriter *>(writer.get())) name = g->filename(iteration);
    if (auto *a = dynamic_cast<AsciiFileDensityGridWriter *>(writer.get())) name = a->filename(iteration);
    strncpy(filename, name.c_str(), nfilename - 1);
    filename[nfilename - 1] = 0;
  });
}
/* a Gadget-style SPH snapshot from particle arrays through host/HDF5Writer.hpp (test input for GadgetSnapshot):
 * periodic < 0 leaves /RuntimePars out, unit_length_in_cgs == 0 leaves /Units out; xH may be null */
int cmih_write_particle_snapshot(const char *filename, int64_t N, const double *pos, const double *m, const double *h,
                                 const double *rho, const double *T, const double *xH, int periodic, const double *boxsize,
                                 double unit_length_in_cgs, double unit_mass_in_cgs, double unit_temperature_in_cgs,
                                 double unit_time_in_cgs, double time, const double *sfr, int64_t Nstar, const double *star_pos,
                                 const double *star_formation_time, const double *star_mass) {
  CMIH_TRY({
    hdf5::HDF5File file;
    hdf5::Group &header = file.root().create_group("Header");
    header.write_attribute("BoxSize", std::array<double, 3>{boxsize[0], boxsize[1], boxsize[2]});
    std::vector<uint32_t> numpart(6, 0);
    numpart[0] = (uint32_t)N;
    numpart[4] = (uint32_t)Nstar;
    header.write_attribute("NumPart_ThisFile", numpart);
    header.write_attribute("Time", time);
    if (periodic >= 0) file.root().create_group("RuntimePars").write_attribute("PeriodicBoundariesOn", int32_t(periodic));
    if (unit_length_in_cgs != 0.) {
      hdf5::Group &units = file.root().create_group("Units");
      units.write_attribute("Unit length in cgs (U_L)", unit_length_in_cgs);
      units.write_attribute("Unit mass in cgs (U_M)", unit_mass_in_cgs);
      units.write_attribute("Unit temperature in cgs (U_T)", unit_temperature_in_cgs);
      units.write_attribute("Unit time in cgs (U_t)", unit_time_in_cgs);
    }
    hdf5::Group &gas = file.root().create_group("PartType0");
    const uint64_t n = (uint64_t)N;
    gas.create_dataset("Coordinates", hdf5::Type::F64, {n, 3}, pos);
    gas.create_dataset("Masses", hdf5::Type::F64, {n}, m);
    gas.create_dataset("SmoothingLength", hdf5::Type::F64, {n}, h);
    gas.create_dataset("Density", hdf5::Type::F64, {n}, rho);
    if (T) gas.create_dataset("Temperature", hdf5::Type::F64, {n}, T);
    if (xH) gas.create_dataset("NeutralFractionH", hdf5::Type::F64, {n}, xH);
    if (sfr) gas.create_dataset("StarFormationRate", hdf5::Type::F64, {n}, sfr);
    if (Nstar > 0) {
      hdf5::Group &stars = file.root().create_group("PartType4");
      const uint64_t ns = (uint64_t)Nstar;
      stars.create_dataset("Coordinates", hdf5::Type::F64, {ns, 3}, star_pos);
      stars.create_dataset("FormationTime", hdf5::Type::F64, {ns}, star_formation_time);
      stars.create_dataset("Masses", hdf5::Type::F64, {ns}, star_mass);
    }
    file.write(filename);
  });
}
/* host/HDF5Reader.hpp probes: a dataset as doubles (returns its element count; shape in dims[0..ndim)) ... */
int cmih_hdf5_dataset(const char *filename, const char *path, double *out, int64_t capacity, int64_t *dims, int *ndim,
                      int64_t *count) {
  CMIH_TRY({
    hdf5::HDF5Input file(filename);
    std::vector<uint64_t> shape;
    const std::vector<double> v = file.read_dataset(path, &shape);
    *count = (int64_t)v.size();
    *ndim = (int)shape.size();
    for (size_t k = 0; k < shape.size() && k < 4; ++k) dims[k] = (int64_t)shape[k];
    if ((int64_t)v.size() <= capacity) std::copy(v.begin(), v.end(), out);
  });
}
/* ... the attribute names of an object, one per line ... */
int cmih_hdf5_attribute_names(const char *filename, const char *path, char *out, int n) {
  CMIH_TRY({
    hdf5::HDF5Input file(filename);
    std::string all;
    for (const std::string &name : file.get_attribute_names(path)) all += name + "\n";
    strncpy(out, all.c_str(), n - 1);
    out[n - 1] = 0;
  });
}
/* ... and one attribute: kind 0 = string -> text, kind 1 = numbers -> values[0..count) */
int cmih_hdf5_attribute(const char *filename, const char *path, const char *name, int kind, char *text, int ntext,
                        double *values, int capacity, int *count) {
  CMIH_TRY({
    hdf5::HDF5Input file(filename);
    if (kind == 0) {
      const std::string v = file.read_string_attribute(path, name);
      strncpy(text, v.c_str(), ntext - 1);
      text[ntext - 1] = 0;
    } else {
      const std::vector<double> v = file.read_double_attribute(path, name);
      *count = (int)v.size();
      for (int k = 0; k < *count && k < capacity; ++k) values[k] = v[k];
    }
  });
}
int cmih_hdf5_exists(const char *filename, const char *path, int *out) {
  CMIH_TRY({
    hdf5::HDF5Input file(filename);
    *out = file.exists(path) ? 1 : 0;
  });
}
/* as cmih_initial_number_density, with the temperature and the neutral fraction of hydrogen of every cell */
int cmih_initial_grid(void *h, int64_t n, double *dens, double *temp, double *xH) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    std::unique_ptr<DensityFunction> f(DensityFunctionFactory::generate(p));
    std::unique_ptr<FractalDensityMask> mask(DensityMaskFactory::generate(p));
    const SimulationBox box(p);
    CartesianCells cells(box, p.get_value<std::array<int32_t, 3>>("DensityGrid:number of cells", {64, 64, 64}));
    if ((int64_t)cells.get_number_of_cells() != n) throw std::runtime_error("wrong number of cells");
    f->initialize();
    cells.set_densities(*f);
    if (mask) {
      mask->initialize();
      mask->apply(cells);
    }
    for (int64_t i = 0; i < n; ++i) {
      dens[i] = cells.number_density[i];
      temp[i] = cells.temperature[i];
      xH[i] = cells.ionic_fraction[i];
    }
  });
}
/* AbundanceModelFactory on the parameter file -> He C N O Ne S relative to H */
int cmih_abundances(void *h, double *out) {
  CMIH_TRY({
    const Abundances a = Abundances::generate(*static_cast<ParameterFile *>(h));
    for (int i = 0; i < CMIB_NUM_ELEMENTS; ++i) out[i] = a.abundance[i];
  });
}
/* CrossSectionsFactory on the parameter file for the types that are plain parameters (FixedValue,
 * Bimodal): sigma[n][14] at the n frequencies; Verner is evaluated on the device (cmib_eval_cross_sections) */
int cmih_parameter_cross_sections(void *h, int64_t n, const double *nu, double *sigma) {
  CMIH_TRY({
    std::unique_ptr<CrossSections> c(CrossSections::generate(*static_cast<ParameterFile *>(h)));
    if (c->kind == CMIB_CROSS_SECTIONS_VERNER) throw std::runtime_error("Verner cross sections live on the device");
    for (int64_t i = 0; i < n; ++i)
      for (int k = 0; k < CMIB_NUM_IONS; ++k)
        sigma[i * CMIB_NUM_IONS + k] = (c->kind == 2 && !(nu[i] < c->frequency_limit)) ? c->high[k] : c->fixed[k];
  });
}
/* n frequencies sampled on the HOST from the parameter file's spectrum for `role` with RandomGenerator(seed):
 * the samplers of csrc/source.cuh driven by the reference's stream (used to build Masked spectra) */
int cmih_sample_spectrum(void *h, const char *role, int32_t seed, int64_t n, double *nu) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    std::unique_ptr<PhotonSourceSpectrum> s(PhotonSourceSpectrum::generate(role, p));
    if (!s) throw std::runtime_error("no spectrum (type None)");
    RandomGenerator rg(seed);
    for (int64_t i = 0; i < n; ++i) nu[i] = s->sample(rg);
  });
}
/* n deviates of the host-side RandomGenerator (RANLUX level 2, host/RandomGenerator.hpp) */
int cmih_random_stream(int32_t seed, int64_t n, double *out) {
  CMIH_TRY({
    RandomGenerator rg(seed);
    for (int64_t i = 0; i < n; ++i) out[i] = rg.get_uniform_random_double();
  });
}
/* test hook for the multi-GPU driver's failure handling (IonizationSimulation.hpp Rendezvous): `nthreads` threads go
 * through `rounds` iterations; thread `failing_thread` reports a failure in iteration `failing_round` (-1: nobody fails).
 * out[round] = 1 if every thread saw the round succeed, 0 if every thread saw it fail, -1 if they disagree.  Returns
 * (does not hang) in every case. */
int cmih_test_rendezvous(int nthreads, int rounds, int failing_thread, int failing_round, int *out) {
  CMIH_TRY({
    Rendezvous rv(nthreads);
    std::vector<std::vector<int>> seen(nthreads, std::vector<int>(rounds, -1));
    for (int r = 0; r < rounds; ++r) { /* as the driver does it: reset, one thread per device, join */
      rv.reset();
      std::vector<std::thread> threads;
      for (int t = 0; t < nthreads; ++t)
        threads.emplace_back([&, t, r]() { seen[t][r] = rv.arrive(!(t == failing_thread && r == failing_round)) ? 1 : 0; });
      for (auto &th : threads) th.join();
    }
    for (int r = 0; r < rounds; ++r) {
      out[r] = seen[0][r];
      for (int t = 1; t < nthreads; ++t)
        if (seen[t][r] != seen[0][r]) out[r] = -1;
    }
  });
}
/* the parameter file's PhotonSourceDistribution: info = {number of sources, total luminosity};
 * positions[capacity][3], weights[capacity] */
int cmih_photon_source_distribution(void *h, double *info, double *positions, double *weights, int capacity) {
  CMIH_TRY({
    ParameterFile &p = *static_cast<ParameterFile *>(h);
    std::unique_ptr<PhotonSourceDistribution> d(PhotonSourceDistributionFactory::generate(p));
    if (!d) throw std::runtime_error("no distribution (type None)");
    const size_t n = d->get_number_of_sources();
    info[0] = (double)n;
    info[1] = d->get_total_luminosity();
    if ((int)n > capacity) throw std::runtime_error("capacity too small");
    for (size_t i = 0; i < n; ++i) {
      const Vec3 x = d->get_position(i);
      positions[3 * i] = x[0]; positions[3 * i + 1] = x[1]; positions[3 * i + 2] = x[2];
      weights[i] = d->get_weight(i);
    }
  });
}

/* ---- the reference's coarse C ABI (c/cmi_c_library.h:31-56, src/CMILibrary.cpp:48-208) ----------
 * Same names, same arguments, same global-singleton semantics, errors abort like cmac_error.  num_thread
 * is ignored (the work runs on the GPU of CMIB_DEVICE, default 0); mapping types "M_over_V" and
 * "centroid" (SPHArrayInterface.hpp). */
} /* extern "C" */

namespace {
IonizationSimulation *global_ionization_simulation = nullptr;
SPHArrayInterface *global_interface = nullptr;
Log *global_log = nullptr;

int cmi_device() {
  const char *e = getenv("CMIB_DEVICE");
  return e ? atoi(e) : 0;
}
template <class F> void cmi_guard(const char *what, F f) {
  try {
    f();
  } catch (const std::exception &e) {
    fprintf(stderr, "%s: %s\n", what, e.what());
    abort();
  }
}
void cmi_init_common(const char *parameter_file, int talk) {
  if (talk) global_log = new Log();
  global_ionization_simulation =
      new IonizationSimulation(true, false, false, -1, parameter_file, cmi_device(), global_log);
}
template <typename TX, typename TH, typename TN>
void cmi_compute(const TX *x, const TX *y, const TX *z, const TH *h, const TH *m, TN *nH, size_t N) {
  cmi_guard("cmi_compute_neutral_fraction", [&]() {
    if (!global_interface || !global_ionization_simulation) throw std::runtime_error("cmi_init has not been called");
    global_interface->reset(x, y, z, h, m, N);
    global_ionization_simulation->initialize(global_interface);
    global_ionization_simulation->run([&](CartesianDensityGrid &grid) { global_interface->write(grid); });
    global_interface->fill_array(nH);
  });
}
} // namespace

extern "C" {

void cmi_init(const char *parameter_file, const int num_thread, const double unit_length_in_SI,
              const double unit_mass_in_SI, const char *mapping_type, const int talk) {
  (void)num_thread;
  cmi_guard("cmi_init", [&]() {
    cmi_init_common(parameter_file, talk);
    global_interface = new SPHArrayInterface(unit_length_in_SI, unit_mass_in_SI, mapping_type);
  });
}
void cmi_init_periodic_dp(const char *parameter_file, const int num_thread, const double unit_length_in_SI,
                          const double unit_mass_in_SI, const double *box_anchor, const double *box_sides,
                          const char *mapping_type, const int talk) {
  (void)num_thread;
  cmi_guard("cmi_init_periodic_dp", [&]() {
    cmi_init_common(parameter_file, talk);
    global_interface = new SPHArrayInterface(unit_length_in_SI, unit_mass_in_SI, box_anchor, box_sides, mapping_type);
  });
}
void cmi_init_periodic_sp(const char *parameter_file, const int num_thread, const double unit_length_in_SI,
                          const double unit_mass_in_SI, const float *box_anchor, const float *box_sides,
                          const char *mapping_type, const int talk) {
  (void)num_thread;
  cmi_guard("cmi_init_periodic_sp", [&]() {
    cmi_init_common(parameter_file, talk);
    global_interface = new SPHArrayInterface(unit_length_in_SI, unit_mass_in_SI, box_anchor, box_sides, mapping_type);
  });
}
void cmi_destroy() {
  delete global_ionization_simulation;
  delete global_interface;
  delete global_log;
  global_ionization_simulation = nullptr;
  global_interface = nullptr;
  global_log = nullptr;
}
void cmi_compute_neutral_fraction_dp(const double *x, const double *y, const double *z, const double *h,
                                     const double *m, double *nH, const size_t N) {
  cmi_compute(x, y, z, h, m, nH, N);
}
void cmi_compute_neutral_fraction_mp(const double *x, const double *y, const double *z, const float *h,
                                     const float *m, float *nH, const size_t N) {
  cmi_compute(x, y, z, h, m, nH, N);
}
void cmi_compute_neutral_fraction_sp(const float *x, const float *y, const float *z, const float *h, const float *m,
                                     float *nH, const size_t N) {
  cmi_compute(x, y, z, h, m, nH, N);
}

/* the device-free half of the SPH coupling, for the CPU test tier: densities the interface maps onto the
 * parameter file's Cartesian grid, and the inverse mapping of a given neutral-fraction field.
 * periodic: box_anchor / box_sides given (else NULL).  dens[ncell] out, xH_cells[ncell] in, nH[N] out. */
int cmih_sph_mapping(const char *paramfile, const char *mapping_type, const double *box_anchor, const double *box_sides,
                     int64_t N, const double *x, const double *y, const double *z, const double *h, const double *m,
                     int64_t ncell, double *dens, const double *xH_cells, double *nH) {
  CMIH_TRY({
    ParameterFile p(paramfile);
    const SimulationBox box(p);
    CartesianCells cells(box, p.get_value<std::array<int32_t, 3>>("DensityGrid:number of cells", {64, 64, 64}));
    if ((int64_t)cells.get_number_of_cells() != ncell) throw std::runtime_error("wrong number of cells");
    std::unique_ptr<SPHArrayInterface> sph(box_anchor ? new SPHArrayInterface(1., 1., box_anchor, box_sides, mapping_type)
                                                      : new SPHArrayInterface(1., 1., mapping_type));
    sph->reset(x, y, z, h, m, (size_t)N);
    cells.set_densities(*sph);
    for (int64_t i = 0; i < ncell; ++i) dens[i] = cells.number_density[i];
    for (int64_t i = 0; i < ncell; ++i) cells.ionic_fraction[i] = xH_cells[i];
    sph->write(cells);
    sph->fill_array(nH);
  });
}

} /* extern "C" */
