/*
 * IonizationSimulation.hpp — C++ host layer of the B200 backend: the reference's
 * plugin classes for the photoionization path, re-implemented as thin owners of
 * parameters that configure one `cmib_context` (include/cmib.h) per GPU.
 *
 * Same class names, parameter keys, defaults and error messages as the reference so
 * that an existing parameter file and an existing caller keep working:
 *
 * The classes live in per-family headers next to this one: HostCommon.hpp (Log, names, SimulationBox),
 * DensityFunctions.hpp, PhotonSourceDistributions.hpp, DevicePlugins.hpp (spectra, cross sections, rates, abundances,
 * re-emission, temperature parameters), DensityGrid.hpp (cells, grid, mask), DensityGridWriters.hpp, HDF5Writer.hpp /
 * HDF5Reader.hpp, SPHArrayInterface.hpp; this file holds the driver.
 *
 *   host layer                      reference (under /root/reference/src)
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
#include "HostCommon.hpp"
#include "DensityFunctions.hpp"
#include "PhotonSourceDistributions.hpp"
#include "DevicePlugins.hpp"
#include "DensityGrid.hpp"
#include "DensityGridWriters.hpp"

namespace cmi {

/* threads of a multi-GPU iteration meet here before they enter a collective: if any of them failed (a shoot
 * error, an out-of-memory in a queue resize), ALL skip the collective — a rank that entered an NCCL call alone
 * would wait for the missing one forever and the run would hang instead of printing the error */
class Rendezvous {
public:
  explicit Rendezvous(int n) : n_(n) {}
  void reset() { failed_ = false; }
  /* returns true when every thread arrived with ok == true */
  bool arrive(bool ok) {
    std::unique_lock<std::mutex> lock(m_);
    if (!ok) failed_ = true;
    const uint64_t generation = generation_;
    if (++count_ == n_) {
      count_ = 0;
      ++generation_;
      cv_.notify_all();
    } else {
      cv_.wait(lock, [&] { return generation_ != generation; });
    }
    return !failed_;
  }

private:
  std::mutex m_;
  std::condition_variable cv_;
  int n_, count_ = 0;
  uint64_t generation_ = 0;
  bool failed_ = false;
};

/* ---- the driver ---- */
class IonizationSimulation {
public:
  /* same leading arguments as the reference (IonizationSimulation.hpp:196-202); num_thread is
   * accepted and ignored (the parallelism is the GPU's); the MPICommunicator* is replaced by the
   * list of devices of this node.  With more than one device the iteration follows the reference's MPI
   * decomposition (IonizationSimulation.cpp:392-397, 458-618) with the collectives of include/cmib.h
   * (cmib_comm_*: NCCL on each context's stream): packets split by global id (MPICommunicator::distribute),
   * every device holds the whole grid, the accumulators are all-reduced, every device updates the cell chunks it owns
   * (cmib_owned_cell), and the opacity records are gathered back. */
  /* task_based = true: the parameter surface of the reference's other driver (`CMacIonize --task-based`,
   * TaskBasedIonizationSimulation.cpp:190-370) on the same GPU path: Monte Carlo parameters come from the
   * `TaskBasedIonizationSimulation:` block (number of iterations 10, number of photons 1e6, random seed,
   * output folder, diffuse field), the periodicity from `DensitySubGridCreator:periodicity`, and the
   * diffuse re-emission handler exists only when `diffuse field` is true (:338-346).  The task queues,
   * buffers, subgrids and source copies of that driver are CPU scheduling and have no counterpart here
   * (their keys are read so that they show up in the used-values file).  The packet conventions of that
   * driver are switched on on the device (cmib_set_packet_conventions: abundance-weighted cross sections
   * carried by the packet and divided out again before the state update, TaskBasedIonizationSimulation.cpp
   * :932-951; heating relative to the hard-coded 3.288e15 / 5.948e15 Hz, DensitySubGrid.hpp:593-612; A_He = 1
   * in the re-emission decision).  What stays as in IonizationSimulation: the arithmetic of the walk
   * (whole-box coordinates instead of subgrid-relative ones: the same cells, positions equal to rounding) and
   * the order of a packet's random draws (its own counter-based stream either way).  Pinned on the reference's
   * own task-based run of lexingtonHII20.param (tests/test_gpu_benchmarks.py). */
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
      CMIB_CALL(cmib_set_packet_conventions(ctx, task_based ? CMIB_CONVENTIONS_TASK_BASED : CMIB_CONVENTIONS_IONIZATION_SIMULATION));
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
      std::vector<cmib_context *> ctxs;
      for (auto &grid : density_grids_) ctxs.push_back(grid->context());
      CMIB_CALL(cmib_comm_init_all(ctxs.data(), (int32_t)ctxs.size()));
      rendezvous_.reset(new Rendezvous((int)ctxs.size()));
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
    for (auto &grid : density_grids_)
      if (grid && grid->context()) cmib_comm_finalize(grid->context());
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
    if (rendezvous_) rendezvous_->reset();
    auto work = [&](size_t d) {
      cmib_context *ctx = density_grids_[d]->context();
      IterationResult &r = part[d];
      bool ok = true;
      auto t0 = clock::now(), t1 = t0;
      try {
        /* MPICommunicator::distribute_block (MPICommunicator.hpp:237-255): contiguous id blocks */
        uint64_t lo, hi;
        cmib_distribute_block((int32_t)d, (int32_t)ndev, 0, numphoton, &lo, &hi);
        density_grids_[d]->reset_grid();
        CMIB_CALL(cmib_update_reemission_probabilities(ctx));
        CMIB_CALL(cmib_synchronize(ctx));
        t0 = clock::now();
        CMIB_CALL(cmib_shoot(ctx, hi - lo, lo, (uint64_t)(int64_t)random_seed_, loop, &r.totweight, r.typecount));
        t1 = clock::now();
      } catch (const std::exception &e) {
        errors[d] = e.what();
        ok = false;
      }
      /* nobody enters the collective unless everybody got here without an error */
      if (rendezvous_ && !rendezvous_->arrive(ok)) return;
      if (!ok) return;
      try {
        /* all-reduce the accumulators, update the owned cell chunks, all-gather the opacity records;
         * totweight is the reduced device-side sum */
        CMIB_CALL(cmib_comm_exchange_and_update(ctx, loop, 0));
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
  /* the host mirror of device 0 is about to be refreshed: collect the per-cell state that stays with the block
   * owners between iterations (metal fractions, heating terms); collective over the devices */
  void gather_state() {
    if (density_grids_.size() < 2) return;
    std::vector<std::string> errors(density_grids_.size());
    std::vector<std::thread> threads;
    for (size_t d = 0; d < density_grids_.size(); ++d)
      threads.emplace_back([&, d] {
        try {
          CMIB_CALL(cmib_comm_gather_state(density_grids_[d]->context()));
          CMIB_CALL(cmib_synchronize(density_grids_[d]->context()));
        } catch (const std::exception &e) {
          errors[d] = e.what();
        }
      });
    for (auto &t : threads) t.join();
    for (const std::string &e : errors)
      if (!e.empty()) throw Error(e);
  }

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
      { gather_state(); grid.download(); density_grid_writer_->write(grid, loop + 1, parameter_file_); }
    }
    gather_state();
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
  std::unique_ptr<Rendezvous> rendezvous_;
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
