/*
 * oracle/ref_harness.cpp — C-ABI probes into the UNMODIFIED CMacIonize reference.
 *
 * TEST INFRASTRUCTURE ONLY: compiled by oracle/build_ref.py together with the
 * reference's own sources (where they lie under /root/reference/src) into
 * oracle/_ref/libcmi_ref.so.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load that library.
 * Nothing here is a restatement: every function below instantiates the
 * reference's own classes and calls the reference's own methods, so the numbers
 * it returns ARE the reference's numbers (parity pinned by construction, and
 * additionally checked against the reference's golden files in tests/).
 *
 * Each probe names the reference entry point it drives.
 */
#include <algorithm>
#include <cfloat>
#include <cinttypes>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <string>
#include <utility>
#include <vector>
#include <dlfcn.h>
#include <omp.h>

/* The probes need to read a few private members of the reference classes
 * (timers, the grid owned by IonizationSimulation, spectrum tables).  GCC does
 * not reorder members on access specifiers, so the object layout is identical to
 * the one the reference translation units were compiled with. */
#define private public
#define protected public
#include "Abundances.hpp"
#include "AbundanceModelFactory.hpp"
#include "CrossSectionsFactory.hpp"
#include "ContinuousPhotonSourceFactory.hpp"
#include "CubicSplineKernel.hpp"
#include "Octree.hpp"
#include "SimulationBox.hpp"
#include "CartesianDensityGrid.hpp"
#include "ChargeTransferRates.hpp"
#include "DistantStarContinuousPhotonSource.hpp"
#include "ExtendedDiscContinuousPhotonSource.hpp"
#include "DensityGridWriter.hpp"
#include "FixedValueCrossSections.hpp"
#include "FixedValueRecombinationRates.hpp"
#include "HeliumLymanContinuumSpectrum.hpp"
#include "HeliumTwoPhotonContinuumSpectrum.hpp"
#include "HomogeneousDensityFunction.hpp"
#include "HydrogenLymanContinuumSpectrum.hpp"
#include "IonizationSimulation.hpp"
#include "IonizationStateCalculator.hpp"
#include "FaucherGiguerePhotonSourceSpectrum.hpp"
#include "IsotropicContinuousPhotonSource.hpp"
#include "UniformPhotonSourceSpectrum.hpp"
#include "LineCoolingData.hpp"
#include "MaskedPhotonSourceSpectrum.hpp"
#include "PhotonSourceSpectrumFactory.hpp"
#include "Photon.hpp"
#include "PhotonSource.hpp"
#include "PlanarContinuousPhotonSource.hpp"
#include "PhotonSourceDistributionFactory.hpp"
#include "PhysicalDiffuseReemissionHandler.hpp"
#include "PlanckPhotonSourceSpectrum.hpp"
#include "RandomGenerator.hpp"
#include "TaskBasedIonizationSimulation.hpp"
#include "SPHArrayInterface.hpp"
#include "SpiralGalaxyContinuousPhotonSource.hpp"
#include "TemperatureCalculator.hpp"
#include "TerminalLog.hpp"
#include "Tracker.hpp"
#include "UnitConverter.hpp"
#include "VernerCrossSections.hpp"
#include "VernerRecombinationRates.hpp"
#undef private
#undef protected

/* ---- data-file resolver used by the generated *DataLocation.hpp headers ---- */
std::string cmi_ref_data_file(const char *name);
/* directory (with trailing slash) of a data set that is addressed by file-name prefix
 * (FaucherGiguereDataLocation.hpp: fg_uvb_dec11/) */
std::string cmi_ref_data_dir(const char *sub, const char *probe_file) {
  const std::string probe = cmi_ref_data_file((std::string(sub) + "/" + probe_file).c_str());
  return probe.substr(0, probe.size() - strlen(probe_file));
}
std::string cmi_ref_data_file(const char *name) {
  const char *env = getenv("CMI_REF_DATA_DIR");
  std::string dir;
  if (env) {
    dir = env;
  } else {
    Dl_info info;
    if (dladdr((void *)&cmi_ref_data_file, &info) && info.dli_fname) {
      std::string so(info.dli_fname);
      size_t p = so.find_last_of('/');
      dir = (p == std::string::npos ? std::string(".") : so.substr(0, p)) + "/data";
    } else {
      dir = "data";
    }
  }
  std::string full = dir + "/" + name;
  std::ifstream probe(full);
  if (!probe.good()) {
    fprintf(stderr, "cmi_ref: data file %s not found\n", full.c_str());
    abort();
  }
  return full;
}

namespace {

/* lazily constructed reference singletons (read-only after construction) */
const VernerCrossSections &verner_xs() {
  static VernerCrossSections xs;
  return xs;
}
const VernerRecombinationRates &verner_rr() {
  static VernerRecombinationRates rr;
  return rr;
}
const ChargeTransferRates &ctr() {
  static ChargeTransferRates c;
  return c;
}
const LineCoolingData &lcd() {
  static LineCoolingData l;
  return l;
}

/* Tracker that records the order in which cells receive update_integrals()
 * calls (DensityGrid.hpp:188-191 calls count_photon under the cell lock). */
struct CellTrace {
  std::vector<int64_t> *sink;
};
class CellOrderTracker : public Tracker {
public:
  int64_t _cell;
  std::vector<int64_t> *_sink;
  CellOrderTracker(int64_t cell, std::vector<int64_t> *sink) : _cell(cell), _sink(sink) {}
  virtual Tracker *duplicate() { return new CellOrderTracker(_cell, _sink); }
  virtual void merge(Tracker *) {}
  virtual void count_photon(const Photon &) { _sink->push_back(_cell); }
  virtual void count_photon(const PhotonPacket &, const double *) {}
  virtual void output_tracker(const std::string) const {}
};

} // namespace

extern "C" {

int cmi_ref_abi_version(void) { return 1; }

int cmi_ref_num_ions(void) { return NUMBER_OF_IONNAMES; }

int cmi_ref_max_threads(void) { return omp_get_max_threads(); }

/* ---------------------------------------------------------------------------
 * a6: VernerCrossSections::get_cross_section (VernerCrossSections.cpp:259-322)
 * sigma is [n][14] row-major (m^2); nu in Hz.
 * ------------------------------------------------------------------------- */
void cmi_ref_verner_cross_sections(int64_t n, const double *nu, double *sigma) {
  const VernerCrossSections &xs = verner_xs();
  for (int64_t i = 0; i < n; ++i)
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion)
      sigma[i * NUMBER_OF_IONNAMES + ion] = xs.get_cross_section(ion, nu[i]);
}

/* a14: VernerRecombinationRates::get_recombination_rate
 * (VernerRecombinationRates.cpp:157-333); alpha is [n][14] (m^3 s^-1). */
void cmi_ref_verner_recombination_rates(int64_t n, const double *T, double *alpha) {
  const VernerRecombinationRates &rr = verner_rr();
  for (int64_t i = 0; i < n; ++i)
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion)
      alpha[i * NUMBER_OF_IONNAMES + ion] = rr.get_recombination_rate(ion, T[i]);
}

/* ---------------------------------------------------------------------------
 * a7-a9: CartesianDensityGrid::interact (CartesianDensityGrid.cpp:375-452) with
 * DensityGrid::get_optical_depth / update_integrals (DensityGrid.hpp:117-197).
 *
 * Cells are given as SoA arrays in the reference's long-index order
 * (ix*ny*nz + iy*nz + iz).  Packets are given explicitly: position,
 * direction, the 14 cross sections, the abundance-corrected He cross section,
 * frequency, weight and the target optical depth.  Packets are processed
 * serially in order so the accumulation order is deterministic.
 *
 * Outputs: J [14][ncell], heat [2][ncell] (accumulated on top of the given
 * contents when accumulate != 0, else zeroed first); final position [np][3];
 * final cell long index (-1 when the packet left the box); number of cells
 * visited with n > 0 (nsteps); optional per-packet cell trace: trace_cells is
 * [np][max_trace], filled with the visit order (only cells with n > 0 are
 * reported because the hook sits inside update_integrals), -1 padded.
 * Returns 0.
 * ------------------------------------------------------------------------- */
int cmi_ref_interact(const double *anchor, const double *sides, const int32_t *ncell,
                     const int32_t *periodic, const double *cell_n, const double *cell_xH,
                     const double *cell_xHe, int64_t np, const double *pos, const double *dir,
                     const double *sigma, const double *sigma_He_corr, const double *nu,
                     const double *weight, const double *tau, int accumulate, double *J,
                     double *heat, double *final_pos, int64_t *final_cell, int32_t *nsteps,
                     int32_t max_trace, int64_t *trace_cells) {
  Box<> box(CoordinateVector<>(anchor[0], anchor[1], anchor[2]),
            CoordinateVector<>(sides[0], sides[1], sides[2]));
  CartesianDensityGrid grid(box, CoordinateVector<int_fast32_t>(ncell[0], ncell[1], ncell[2]),
                            CoordinateVector<bool>(periodic[0] != 0, periodic[1] != 0,
                                                   periodic[2] != 0),
                            false, nullptr);
  const int64_t nc = (int64_t)ncell[0] * ncell[1] * ncell[2];
  std::vector<int64_t> sink;
  std::vector<CellOrderTracker *> trackers;
  const bool want_steps = (nsteps != nullptr) || (max_trace > 0 && trace_cells != nullptr);
  for (int64_t i = 0; i < nc; ++i) {
    IonizationVariables &iv = grid._ionization_variables[i];
    iv.set_number_density(cell_n[i]);
    iv.set_ionic_fraction(ION_H_n, cell_xH[i]);
    iv.set_ionic_fraction(ION_He_n, cell_xHe ? cell_xHe[i] : 0.);
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion)
      iv.set_mean_intensity(ion, accumulate ? J[ion * nc + i] : 0.);
    iv.set_heating(HEATINGTERM_H, accumulate ? heat[i] : 0.);
    iv.set_heating(HEATINGTERM_He, accumulate ? heat[nc + i] : 0.);
    if (want_steps) {
      trackers.push_back(new CellOrderTracker(i, &sink));
      iv.add_tracker(trackers.back());
    }
  }
  for (int64_t p = 0; p < np; ++p) {
    Photon photon(CoordinateVector<>(pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]),
                  CoordinateVector<>(dir[3 * p], dir[3 * p + 1], dir[3 * p + 2]), nu[p]);
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion)
      photon.set_cross_section(ion, sigma[p * NUMBER_OF_IONNAMES + ion]);
    photon.set_cross_section_He_corr(sigma_He_corr[p]);
    photon.set_weight(weight[p]);
    sink.clear();
    DensityGrid::iterator it = grid.interact(photon, tau[p]);
    const CoordinateVector<> fp = photon.get_position();
    final_pos[3 * p] = fp.x();
    final_pos[3 * p + 1] = fp.y();
    final_pos[3 * p + 2] = fp.z();
    final_cell[p] = (it == grid.end()) ? -1 : (int64_t)it.get_index();
    if (nsteps) nsteps[p] = (int32_t)sink.size();
    if (max_trace > 0 && trace_cells) {
      for (int32_t k = 0; k < max_trace; ++k)
        trace_cells[p * max_trace + k] = (k < (int32_t)sink.size()) ? sink[k] : -1;
    }
  }
  for (int64_t i = 0; i < nc; ++i) {
    IonizationVariables &iv = grid._ionization_variables[i];
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) J[ion * nc + i] = iv.get_mean_intensity(ion);
    heat[i] = iv.get_heating(HEATINGTERM_H);
    heat[nc + i] = iv.get_heating(HEATINGTERM_He);
    iv.add_tracker(nullptr); /* do not let ~IonizationVariables free ours twice */
  }
  for (size_t i = 0; i < trackers.size(); ++i) delete trackers[i];
  return 0;
}

/* CartesianDensityGrid::integrate_optical_depth (src/CartesianDensityGrid.cpp:328-363) for np photons
 * on a grid given cell by cell (same arguments as cmi_ref_interact) */
int cmi_ref_integrate_optical_depth(const double *anchor, const double *sides, const int32_t *ncell,
                                    const int32_t *periodic, const double *cell_n, const double *cell_xH,
                                    const double *cell_xHe, int64_t np, const double *pos, const double *dir,
                                    const double *sigma_H, const double *sigma_He_corr, double *optical_depth) {
  Box<> box(CoordinateVector<>(anchor[0], anchor[1], anchor[2]), CoordinateVector<>(sides[0], sides[1], sides[2]));
  CartesianDensityGrid grid(box, CoordinateVector<int_fast32_t>(ncell[0], ncell[1], ncell[2]),
                            CoordinateVector<bool>(periodic[0] != 0, periodic[1] != 0, periodic[2] != 0), false, nullptr);
  const int64_t nc = (int64_t)ncell[0] * ncell[1] * ncell[2];
  for (int64_t i = 0; i < nc; ++i) {
    IonizationVariables &iv = grid._ionization_variables[i];
    iv.set_number_density(cell_n[i]);
    iv.set_ionic_fraction(ION_H_n, cell_xH[i]);
    iv.set_ionic_fraction(ION_He_n, cell_xHe ? cell_xHe[i] : 0.);
  }
  for (int64_t p = 0; p < np; ++p) {
    Photon photon(CoordinateVector<>(pos[3 * p], pos[3 * p + 1], pos[3 * p + 2]),
                  CoordinateVector<>(dir[3 * p], dir[3 * p + 1], dir[3 * p + 2]), 3.3e15);
    photon.set_cross_section(ION_H_n, sigma_H[p]);
    photon.set_cross_section_He_corr(sigma_He_corr[p]);
    optical_depth[p] = grid.integrate_optical_depth(photon);
  }
  return 0;
}

/* ---------------------------------------------------------------------------
 * a1: IonizationSimulation (IonizationSimulation.cpp:101-679) on a parameter
 * file, with an in-process DensityGridWriter that captures every field.
 *
 * fields is [32][ncell]: n, T, x[14], raw J[14] of the last iteration, normalised
 * heat[2]; the caller passes the ncell capacity.
 * Returns number of cells, or <0 on error.  times[0] = "Total photon shooting
 * time" (s), times[1] = "Total cell update time" (s) — the reference's own
 * timers (IonizationSimulation.cpp:667-674).
 * ------------------------------------------------------------------------- */
namespace {
class CaptureWriter : public DensityGridWriter {
public:
  double *_out;
  int64_t _cap;
  int64_t _n;
  CaptureWriter(double *out, int64_t cap)
      : DensityGridWriter(".", false, DensityGridWriterFields(false), nullptr), _out(out),
        _cap(cap), _n(0) {}
  virtual void write(DensityGrid &grid, uint_fast32_t, ParameterFile &, double,
                     const InternalHydroUnits *) {
    _n = grid.get_number_of_cells();
    if (_n > _cap) return;
    for (auto it = grid.begin(); it != grid.end(); ++it) {
      const int64_t i = it.get_index();
      const IonizationVariables &iv = it.get_ionization_variables();
      _out[0 * _n + i] = iv.get_number_density();
      _out[1 * _n + i] = iv.get_temperature();
      for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) {
        _out[(2 + ion) * _n + i] = iv.get_ionic_fraction(ion);
        _out[(16 + ion) * _n + i] = iv.get_mean_intensity(ion);
      }
      _out[30 * _n + i] = iv.get_heating(HEATINGTERM_H);
      _out[31 * _n + i] = iv.get_heating(HEATINGTERM_He);
    }
  }
  /* task-based grid (TaskBasedIonizationSimulation.cpp:1091): the cells of all original subgrids, placed by the
   * cell midpoint into the Cartesian cell order of the whole box */
  virtual void write(DensitySubGridCreator<DensitySubGrid> &creator, const uint_fast32_t,
                     ParameterFile &params, double) {
    const CoordinateVector<> anchor = params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:anchor");
    const CoordinateVector<> sides = params.get_physical_vector<QUANTITY_LENGTH>("SimulationBox:sides");
    const CoordinateVector<int_fast32_t> nc =
        params.get_value<CoordinateVector<int_fast32_t>>("DensityGrid:number of cells");
    _n = (int64_t)nc.x() * nc.y() * nc.z();
    if (_n > _cap) return;
    for (auto gridit = creator.begin(); gridit != creator.original_end(); ++gridit) {
      for (auto cellit = (*gridit).begin(); cellit != (*gridit).end(); ++cellit) {
        const CoordinateVector<> mid = cellit.get_cell_midpoint();
        const int64_t ix = (int64_t)((mid.x() - anchor.x()) / sides.x() * nc.x());
        const int64_t iy = (int64_t)((mid.y() - anchor.y()) / sides.y() * nc.y());
        const int64_t iz = (int64_t)((mid.z() - anchor.z()) / sides.z() * nc.z());
        const int64_t i = (ix * nc.y() + iy) * nc.z() + iz;
        const IonizationVariables &iv = cellit.get_ionization_variables();
        _out[0 * _n + i] = iv.get_number_density();
        _out[1 * _n + i] = iv.get_temperature();
        for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) {
          _out[(2 + ion) * _n + i] = iv.get_ionic_fraction(ion);
          _out[(16 + ion) * _n + i] = iv.get_mean_intensity(ion);
        }
        _out[30 * _n + i] = iv.get_heating(HEATINGTERM_H);
        _out[31 * _n + i] = iv.get_heating(HEATINGTERM_He);
      }
    }
  }
  virtual void write(DensitySubGridCreator<HydroDensitySubGrid> &, const uint_fast32_t,
                     ParameterFile &, double) {}
};
} // namespace

int64_t cmi_ref_run_paramfile(const char *paramfile, int num_threads, int verbose,
                              double *fields, int64_t ncell_capacity, double *times) {
  TerminalLog *log = verbose ? new TerminalLog(LOGLEVEL_STATUS) : nullptr;
  int64_t n = -1;
  {
    IonizationSimulation simulation(false, false, false, num_threads, paramfile, nullptr, log);
    simulation.initialize();
    CaptureWriter writer(fields, ncell_capacity);
    simulation.run(&writer);
    n = writer._n;
    if (times) {
      times[0] = simulation._photon_propagation_timer.value();
      times[1] = simulation._cell_update_timer.value();
    }
  }
  delete log;
  return n;
}

/* f2: the reference's task-based driver (TaskBasedIonizationSimulation.cpp:190-1097: `CMacIonize --task-based`)
 * on a parameter file; fields as cmi_ref_run_paramfile (J and heat: what the last temperature step left, i.e.
 * already divided by the abundances, :932-951).  The driver's own writer (built from the parameter file) is
 * replaced by the capturing one before the run; run() writes through the member, not through its argument (:1091). */
int64_t cmi_ref_run_paramfile_taskbased(const char *paramfile, int num_threads, int verbose, double *fields,
                                        int64_t ncell_capacity) {
  TerminalLog *log = verbose ? new TerminalLog(LOGLEVEL_STATUS) : nullptr;
  int64_t n = -1;
  {
    if (num_threads <= 0) num_threads = omp_get_max_threads();
    TaskBasedIonizationSimulation simulation(num_threads, paramfile, false, false, log);
    simulation.initialize();
    CaptureWriter *writer = new CaptureWriter(fields, ncell_capacity);
    delete simulation._density_grid_writer;
    simulation._density_grid_writer = writer; /* deleted by the simulation's destructor */
    simulation.run(nullptr);
    n = writer->_n;
  }
  delete log;
  return n;
}

/* ---------------------------------------------------------------------------
 * Charge transfer (ChargeTransferRates.cpp:44-395).  out is [n][3][14]:
 * [0] recombination with H, [1] ionization with H, [2] recombination with He;
 * entries the reference refuses to evaluate (H with itself, He with itself)
 * are returned as 0.
 * ------------------------------------------------------------------------- */
void cmi_ref_charge_transfer(int64_t n, const double *T4, double *out) {
  const ChargeTransferRates &c = ctr();
  for (int64_t i = 0; i < n; ++i) {
    double *o = out + i * 3 * NUMBER_OF_IONNAMES;
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) {
      o[ion] = (ion == ION_H_n) ? 0. : c.get_charge_transfer_recombination_rate_H(ion, T4[i]);
      o[NUMBER_OF_IONNAMES + ion] =
          (ion == ION_H_n) ? 0. : c.get_charge_transfer_ionization_rate_H(ion, T4[i]);
      o[2 * NUMBER_OF_IONNAMES + ion] =
          (ion == ION_He_n) ? 0. : c.get_charge_transfer_recombination_rate_He(ion, T4[i]);
    }
  }
}

/* ---------------------------------------------------------------------------
 * LineCoolingData constants (LineCoolingData.cpp:42-1394), flattened:
 *   cs5[10][10][7], A5[10][10], E5[10][10], invw5[10][5],
 *   cs2[3][7], A2[3], E2[3], invw2[3][2], prefactor      (984 doubles)
 * Used once by tools/gen_linecooling_data.py to carry the atomic data.
 * ------------------------------------------------------------------------- */
int cmi_ref_linecooling_tables(double *out, int capacity) {
  const LineCoolingData &l = lcd();
  int k = 0;
  if (capacity < 984) return -1;
  for (int e = 0; e < 10; ++e)
    for (int t = 0; t < 10; ++t)
      for (int c = 0; c < 7; ++c) out[k++] = l._five_level_collision_strength[e][t][c];
  for (int e = 0; e < 10; ++e)
    for (int t = 0; t < 10; ++t) out[k++] = l._five_level_transition_probability[e][t];
  for (int e = 0; e < 10; ++e)
    for (int t = 0; t < 10; ++t) out[k++] = l._five_level_energy_difference[e][t];
  for (int e = 0; e < 10; ++e)
    for (int t = 0; t < 5; ++t) out[k++] = l._five_level_inverse_statistical_weight[e][t];
  for (int e = 0; e < 3; ++e)
    for (int c = 0; c < 7; ++c) out[k++] = l._two_level_collision_strength[e][c];
  for (int e = 0; e < 3; ++e) out[k++] = l._two_level_transition_probability[e];
  for (int e = 0; e < 3; ++e) out[k++] = l._two_level_energy_difference[e];
  for (int e = 0; e < 3; ++e)
    for (int t = 0; t < 2; ++t) out[k++] = l._two_level_inverse_statistical_weight[e][t];
  out[k++] = l._collision_strength_prefactor;
  return k;
}

/* LineCoolingData::get_cooling (LineCoolingData.cpp:1767-1848); abund is [n][13]
 * in LineCoolingData element order (NI NII OI OII OIII NeIII SII SIII CII CIII
 * NIII NeII SIV). */
void cmi_ref_linecooling_get_cooling(int64_t n, const double *T, const double *ne,
                                     const double *abund, double *cooling) {
  const LineCoolingData &l = lcd();
  for (int64_t i = 0; i < n; ++i)
    cooling[i] = l.get_cooling(T[i], ne[i], abund + i * LINECOOLINGDATA_NUMELEMENTS);
}

/* LineCoolingData::solve_system_of_linear_equations (:1492-1555); A is [n][25],
 * B is [n][5], both overwritten; status [n]. */
void cmi_ref_solve5(int64_t n, double *A, double *B, int32_t *status) {
  for (int64_t i = 0; i < n; ++i) {
    double a[5][5];
    memcpy(a, A + 25 * i, sizeof(a));
    status[i] = LineCoolingData::solve_system_of_linear_equations(a, B + 5 * i);
    memcpy(A + 25 * i, a, sizeof(a));
  }
}

/* PhysicalDiffuseReemissionHandler::set_reemission_probabilities
 * (PhysicalDiffuseReemissionHandler.hpp:66-106); out is [n][5]. */
void cmi_ref_reemission_probabilities(int64_t n, const double *T, double *out) {
  static PhysicalDiffuseReemissionHandler *handler = nullptr;
  if (!handler) handler = new PhysicalDiffuseReemissionHandler(verner_xs());
  for (int64_t i = 0; i < n; ++i) {
    IonizationVariables iv;
    iv.set_temperature(T[i]);
    handler->set_reemission_probabilities(iv);
    for (int k = 0; k < NUMBER_OF_REEMISSIONPROBABILITIES; ++k)
      out[i * NUMBER_OF_REEMISSIONPROBABILITIES + k] = iv.get_reemission_probability(k);
  }
}

namespace {
struct RatesHolder {
  RecombinationRates *rates;
  bool owned;
  RatesHolder(int kind, const double *fixed) {
    if (kind == 1) {
      rates = const_cast<VernerRecombinationRates *>(&verner_rr());
      owned = false;
    } else {
      rates = new FixedValueRecombinationRates(fixed[0], fixed[1], fixed[2], fixed[3], fixed[4],
                                               fixed[5], fixed[6], fixed[7], fixed[8], fixed[9],
                                               fixed[10], fixed[11], fixed[12], fixed[13]);
      owned = true;
    }
  }
  ~RatesHolder() {
    if (owned) delete rates;
  }
};
} // namespace

/* ---------------------------------------------------------------------------
 * a13/a14: IonizationStateCalculator::calculate_ionization_state(jfac, hfac,
 * cell) (IonizationStateCalculator.cpp:70-272).  SoA inputs J[14][n],
 * heat[2][n], ndens[n], T[n]; abundances[6] = He C N O Ne S; rr_kind 1 = Verner,
 * 0 = FixedValue(rr_fixed[14]).  Outputs x[14][n], heat_out[2][n].
 * ------------------------------------------------------------------------- */
void cmi_ref_ionization_state(int64_t n, double jfac, double hfac, const double *abundances,
                              int rr_kind, const double *rr_fixed, const double *J,
                              const double *heat, const double *ndens, const double *T,
                              double *x, double *heat_out) {
  Abundances ab(abundances[0], abundances[1], abundances[2], abundances[3], abundances[4],
                abundances[5]);
  RatesHolder rh(rr_kind, rr_fixed);
  IonizationStateCalculator calc(1., ab, *rh.rates, ctr());
  for (int64_t i = 0; i < n; ++i) {
    IonizationVariables iv;
    iv.set_number_density(ndens[i]);
    iv.set_temperature(T[i]);
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) iv.set_mean_intensity(ion, J[ion * n + i]);
    iv.set_heating(HEATINGTERM_H, heat[i]);
    iv.set_heating(HEATINGTERM_He, heat[n + i]);
    calc.calculate_ionization_state(jfac, hfac, iv);
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) x[ion * n + i] = iv.get_ionic_fraction(ion);
    heat_out[i] = iv.get_heating(HEATINGTERM_H);
    heat_out[n + i] = iv.get_heating(HEATINGTERM_He);
  }
}

/* IonizationStateCalculator::compute_ionization_states_hydrogen_helium (:649-753) */
void cmi_ref_h_he_state(int64_t n, const double *alphaH, const double *alphaHe, const double *jH,
                        const double *jHe, const double *nH, const double *AHe, const double *T,
                        double *h0, double *he0) {
  for (int64_t i = 0; i < n; ++i)
    IonizationStateCalculator::compute_ionization_states_hydrogen_helium(
        alphaH[i], alphaHe[i], jH[i], jHe[i], nH[i], AHe[i], T[i], h0[i], he0[i]);
}

/* ---------------------------------------------------------------------------
 * a15: TemperatureCalculator::compute_cooling_and_heating_balance
 * (TemperatureCalculator.cpp:207-501).  j is [n][14] ALREADY normalised, h is
 * [n][2] already normalised.  Outputs h0, he0, gain, loss [n], metals [n][12].
 * ------------------------------------------------------------------------- */
void cmi_ref_cooling_heating_balance(int64_t n, const double *T, const double *ndens,
                                     const double *j, const double *h, const double *abundances,
                                     double pahfac, double crfac, double crscale,
                                     const double *midz, int rr_kind, const double *rr_fixed,
                                     double *h0, double *he0, double *gain, double *loss,
                                     double *metals) {
  Abundances ab(abundances[0], abundances[1], abundances[2], abundances[3], abundances[4],
                abundances[5]);
  RatesHolder rh(rr_kind, rr_fixed);
  for (int64_t i = 0; i < n; ++i) {
    IonizationVariables iv;
    iv.set_number_density(ndens[i]);
    TemperatureCalculator::compute_cooling_and_heating_balance(
        h0[i], he0[i], gain[i], loss[i], T[i], iv,
        CoordinateVector<>(0., 0., midz ? midz[i] : 0.), j + i * NUMBER_OF_IONNAMES, ab,
        h + i * NUMBER_OF_HEATINGTERMS, pahfac, crfac, crscale, lcd(), *rh.rates, ctr());
    for (int m = 0; m < 12; ++m) metals[i * 12 + m] = iv.get_ionic_fraction(2 + m);
  }
}

/* ---------------------------------------------------------------------------
 * a15: TemperatureCalculator::calculate_temperature(cell, jfac, hfac, midpoint)
 * (TemperatureCalculator.cpp:567-931).  SoA inputs J[14][n], heat[2][n],
 * ndens[n], T[n], cr_factor[n] (may be NULL = 1), midz[n] (may be NULL = 0).
 * tparams = {pahfac, crfac, crlim, crscale, minimum_ionized_temperature,
 *            epsilon_convergence, maximum_number_of_iterations}.
 * Outputs T_out[n], x[14][n], heat_out[2][n].
 * ------------------------------------------------------------------------- */
void cmi_ref_temperature(int64_t n, double jfac, double hfac, const double *abundances,
                         int rr_kind, const double *rr_fixed, const double *tparams,
                         const double *J, const double *heat, const double *ndens,
                         const double *T, const double *cr_factor, const double *midz,
                         double *T_out, double *x, double *heat_out) {
  Abundances ab(abundances[0], abundances[1], abundances[2], abundances[3], abundances[4],
                abundances[5]);
  RatesHolder rh(rr_kind, rr_fixed);
  TemperatureCalculator calc(true, 0, 1., ab, tparams[5], (uint_fast32_t)tparams[6], tparams[0],
                             tparams[1], tparams[2], tparams[3], tparams[4], lcd(), *rh.rates,
                             ctr(), nullptr);
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t i = 0; i < n; ++i) {
    IonizationVariables iv;
    iv.set_number_density(ndens[i]);
    iv.set_temperature(T[i]);
    if (cr_factor) iv.set_cosmic_ray_factor(cr_factor[i]);
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) iv.set_mean_intensity(ion, J[ion * n + i]);
    iv.set_heating(HEATINGTERM_H, heat[i]);
    iv.set_heating(HEATINGTERM_He, heat[n + i]);
    calc.calculate_temperature(iv, jfac, hfac, CoordinateVector<>(0., 0., midz ? midz[i] : 0.));
    T_out[i] = iv.get_temperature();
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) x[ion * n + i] = iv.get_ionic_fraction(ion);
    heat_out[i] = iv.get_heating(HEATINGTERM_H);
    heat_out[n + i] = iv.get_heating(HEATINGTERM_He);
  }
}

/* ---------------------------------------------------------------------------
 * Spectrum tables as the reference builds them.
 *  Planck (PlanckPhotonSourceSpectrum.cpp:53-115): out [3][1000] = cdf, logcdf, lognu
 *  H-Lyc / He-Lyc (HydrogenLymanContinuumSpectrum.cpp:40-122,
 *    HeliumLymanContinuumSpectrum.cpp): freq[1000], temp[100], cdf[100][1000];
 *    xs_kind 1 = Verner cross sections, 0 = FixedValue(xs_fixed[14])
 *  He two-photon (HeliumTwoPhotonContinuumSpectrum.cpp:44-101): freq[1000], cdf[1000]
 * ------------------------------------------------------------------------- */
void cmi_ref_planck_tables(double temperature, double *out) {
  PlanckPhotonSourceSpectrum sp(temperature, -1., nullptr);
  for (int i = 0; i < 1000; ++i) {
    out[i] = sp._cumulative_distribution[i];
    out[1000 + i] = sp._log_cumulative_distribution[i];
    out[2000 + i] = sp._log_frequency[i];
  }
}

namespace {
CrossSections *make_xs(int xs_kind, const double *f) {
  if (xs_kind == 1) return new VernerCrossSections();
  return new FixedValueCrossSections(f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], f[9],
                                     f[10], f[11], f[12], f[13]);
}
} // namespace

void cmi_ref_lyc_tables(int which /*0 = H, 1 = He*/, int xs_kind, const double *xs_fixed,
                        double *freq, double *temp, double *cdf) {
  CrossSections *xs = make_xs(xs_kind, xs_fixed);
  if (which == 0) {
    HydrogenLymanContinuumSpectrum sp(*xs);
    for (int i = 0; i < 1000; ++i) freq[i] = sp._frequency[i];
    for (int t = 0; t < 100; ++t) {
      temp[t] = sp._temperature[t];
      for (int i = 0; i < 1000; ++i) cdf[t * 1000 + i] = sp._cumulative_distribution[t][i];
    }
  } else {
    HeliumLymanContinuumSpectrum sp(*xs);
    for (int i = 0; i < 1000; ++i) freq[i] = sp._frequency[i];
    for (int t = 0; t < 100; ++t) {
      temp[t] = sp._temperature[t];
      for (int i = 0; i < 1000; ++i) cdf[t * 1000 + i] = sp._cumulative_distribution[t][i];
    }
  }
  delete xs;
}

void cmi_ref_he2pc_tables(double *freq, double *cdf) {
  HeliumTwoPhotonContinuumSpectrum sp;
  for (int i = 0; i < 1000; ++i) {
    freq[i] = sp._frequency[i];
    cdf[i] = sp._cumulative_distribution[i];
  }
}

/* Sample n frequencies from a reference spectrum with the reference's own RNG
 * (statistical parity of the samplers).  which: 0 Planck(T), 1 H-Lyc(T) Verner,
 * 2 He-Lyc(T) Verner, 3 He two-photon. */
void cmi_ref_sample_spectrum(int which, double temperature, int seed, int64_t n, double *nu) {
  RandomGenerator rg(seed);
  if (which == 0) {
    PlanckPhotonSourceSpectrum sp(temperature, -1., nullptr);
    for (int64_t i = 0; i < n; ++i) nu[i] = sp.get_random_frequency(rg, 0.);
  } else if (which == 1) {
    HydrogenLymanContinuumSpectrum sp(verner_xs());
    for (int64_t i = 0; i < n; ++i) nu[i] = sp.get_random_frequency(rg, temperature);
  } else if (which == 2) {
    HeliumLymanContinuumSpectrum sp(verner_xs());
    for (int64_t i = 0; i < n; ++i) nu[i] = sp.get_random_frequency(rg, temperature);
  } else {
    HeliumTwoPhotonContinuumSpectrum sp;
    for (int64_t i = 0; i < n; ++i) nu[i] = sp.get_random_frequency(rg, temperature);
  }
}

/* IsotropicContinuousPhotonSource::get_random_incoming_direction (src/IsotropicContinuousPhotonSource.hpp:106-180)
 * n times with RandomGenerator(seed); `uniforms` receives the five deviates each call consumed
 * (a second generator with the same seed replays the stream), so that a re-implementation can be
 * fed exactly the same numbers. */
void cmi_ref_isotropic_incoming(const double *anchor, const double *sides, int seed, int64_t n,
                                double *uniforms, double *pos, double *dir) {
  const Box<> box(CoordinateVector<>(anchor[0], anchor[1], anchor[2]),
                  CoordinateVector<>(sides[0], sides[1], sides[2]));
  IsotropicContinuousPhotonSource source(box);
  RandomGenerator rg(seed), replay(seed);
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < 5; ++k) uniforms[5 * i + k] = replay.get_uniform_random_double();
    const std::pair<CoordinateVector<>, CoordinateVector<>> pd = source.get_random_incoming_direction(rg);
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = pd.first[k];
      dir[3 * i + k] = pd.second[k];
    }
  }
}

/* PlanarContinuousPhotonSource::get_random_incoming_direction (src/PlanarContinuousPhotonSource.hpp:179-204),
 * n times, with the four deviates each call consumed */
void cmi_ref_planar_incoming(int axis, double intercept, const double *anchor, const double *sides, int seed,
                             int64_t n, double *uniforms, double *pos, double *dir) {
  const char *names[3] = {"x", "y", "z"};
  PlanarContinuousPhotonSource source(names[axis], intercept, anchor[0], anchor[1], sides[0], sides[1], 1.e48);
  RandomGenerator rg(seed), replay(seed);
  for (int64_t i = 0; i < n; ++i) {
    for (int k = 0; k < 4; ++k) uniforms[4 * i + k] = replay.get_uniform_random_double();
    const std::pair<CoordinateVector<>, CoordinateVector<>> pd = source.get_random_incoming_direction(rg);
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = pd.first[k];
      dir[3 * i + k] = pd.second[k];
    }
  }
}

/* DistantStarContinuousPhotonSource::get_random_incoming_direction (src/DistantStarContinuousPhotonSource.hpp:164-192)
 * n times with RandomGenerator(seed); also the exposed surface area */
double cmi_ref_distant_star_incoming(const double *anchor, const double *sides, const double *star, int seed, int64_t n,
                                     double *pos, double *dir) {
  const Box<> box(CoordinateVector<>(anchor[0], anchor[1], anchor[2]), CoordinateVector<>(sides[0], sides[1], sides[2]));
  DistantStarContinuousPhotonSource source(CoordinateVector<>(star[0], star[1], star[2]), box);
  RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i) {
    const std::pair<CoordinateVector<>, CoordinateVector<>> pd = source.get_random_incoming_direction(rg);
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = pd.first[k];
      dir[3 * i + k] = pd.second[k];
    }
  }
  return source.get_total_surface_area();
}

/* ExtendedDiscContinuousPhotonSource::get_random_incoming_direction (src/ExtendedDiscContinuousPhotonSource.hpp:148-197)
 * n times with RandomGenerator(seed) */
void cmi_ref_extended_disc_incoming(const double *anchor, const double *sides, const char *axis, double origin,
                                    double scale_height, int seed, int64_t n, double *pos, double *dir) {
  const Box<> box(CoordinateVector<>(anchor[0], anchor[1], anchor[2]), CoordinateVector<>(sides[0], sides[1], sides[2]));
  ExtendedDiscContinuousPhotonSource source(box, axis, origin, scale_height, 1.e48);
  RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i) {
    const std::pair<CoordinateVector<>, CoordinateVector<>> pd = source.get_random_incoming_direction(rg);
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = pd.first[k];
      dir[3 * i + k] = pd.second[k];
    }
  }
}

/* SpiralGalaxyContinuousPhotonSource::get_random_incoming_direction (src/SpiralGalaxyContinuousPhotonSource.hpp:131-190)
 * n times with RandomGenerator(seed) */
void cmi_ref_spiral_galaxy_incoming(const double *anchor, const double *sides, double r_stars, double h_stars,
                                    double B_over_T, int seed, int64_t n, double *pos, double *dir) {
  const Box<> box(CoordinateVector<>(anchor[0], anchor[1], anchor[2]), CoordinateVector<>(sides[0], sides[1], sides[2]));
  SpiralGalaxyContinuousPhotonSource source(box, r_stars, h_stars, B_over_T);
  RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i) {
    const std::pair<CoordinateVector<>, CoordinateVector<>> pd = source.get_random_incoming_direction(rg);
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = pd.first[k];
      dir[3 * i + k] = pd.second[k];
    }
  }
}

/* AbundanceModelFactory::generate on a parameter file -> He C N O Ne S */
void cmi_ref_abundances(const char *paramfile, double *out) {
  ParameterFile params(paramfile);
  AbundanceModel *model = AbundanceModelFactory::generate(params, nullptr);
  const Abundances a = model->get_abundances();
  for (int i = 0; i < NUMBER_OF_ELEMENTNAMES; ++i) out[i] = a.get_abundance(i);
  delete model;
}

/* CrossSectionsFactory::generate on a parameter file, evaluated at n frequencies -> sigma[n][14] */
void cmi_ref_parameter_cross_sections(const char *paramfile, int64_t n, const double *nu, double *sigma) {
  ParameterFile params(paramfile);
  CrossSections *xs = CrossSectionsFactory::generate(params, nullptr);
  for (int64_t i = 0; i < n; ++i)
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) sigma[i * NUMBER_OF_IONNAMES + ion] = xs->get_cross_section(ion, nu[i]);
  delete xs;
}

/* PhotonSourceSpectrumFactory::generate(role, params) of type Masked on a parameter file: its frequency
 * bins, cumulative distribution and total flux; returns the number of bins (or -needed) */
int cmi_ref_masked_spectrum(const char *paramfile, const char *role, double *freq, double *cdf, int capacity,
                            double *total_flux) {
  ParameterFile params(paramfile);
  PhotonSourceSpectrum *sp = PhotonSourceSpectrumFactory::generate(role, params, nullptr);
  MaskedPhotonSourceSpectrum *m = dynamic_cast<MaskedPhotonSourceSpectrum *>(sp);
  if (m == nullptr) return 0;
  const int n = (int)m->_frequency_bins.size();
  if (n <= capacity) {
    for (int i = 0; i < n; ++i) {
      freq[i] = m->_frequency_bins[i];
      cdf[i] = m->_cumulative_distribution[i];
    }
    *total_flux = m->get_total_flux();
  }
  delete sp;
  return n <= capacity ? n : -n;
}

/* A PhotonSource wired from a parameter file exactly as IonizationSimulation does it
 * (src/IonizationSimulation.cpp:150-181). */
namespace {
struct RefPhotonSource {
  ParameterFile params;
  SimulationBox box;
  AbundanceModel *abundance_model;
  Abundances abundances;
  CrossSections *cross_sections;
  PhotonSourceDistribution *distribution;
  PhotonSourceSpectrum *spectrum;
  ContinuousPhotonSource *continuous;
  PhotonSourceSpectrum *continuous_spectrum;
  PhotonSource *source;
  explicit RefPhotonSource(const char *paramfile)
      : params(paramfile), box(params), abundance_model(AbundanceModelFactory::generate(params, nullptr)),
        abundances(abundance_model->get_abundances()), cross_sections(CrossSectionsFactory::generate(params, nullptr)),
        distribution(PhotonSourceDistributionFactory::generate(params, nullptr)),
        spectrum(PhotonSourceSpectrumFactory::generate("PhotonSourceSpectrum", params, nullptr)),
        continuous(ContinuousPhotonSourceFactory::generate(box.get_box(), params, nullptr)),
        continuous_spectrum(PhotonSourceSpectrumFactory::generate("ContinuousPhotonSourceSpectrum", params, nullptr)),
        source(new PhotonSource(distribution, spectrum, continuous, continuous_spectrum, abundances, *cross_sections,
                                params, nullptr)) {}
  ~RefPhotonSource() {
    delete source;
    delete continuous_spectrum;
    delete continuous;
    delete spectrum;
    delete distribution;
    delete cross_sections;
    delete abundance_model;
  }
};
} // namespace

/* n x PhotonSource::get_random_photon (src/PhotonSource.cpp:208-249) with RandomGenerator(seed) */
void cmi_ref_random_photons(const char *paramfile, int seed, int64_t n, double *pos, double *dir, double *nu,
                            double *sigma, double *she, double *weight) {
  RefPhotonSource s(paramfile);
  RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i) {
    Photon ph = s.source->get_random_photon(rg);
    for (int k = 0; k < 3; ++k) {
      pos[3 * i + k] = ph.get_position()[k];
      dir[3 * i + k] = ph.get_direction()[k];
    }
    nu[i] = ph.get_energy();
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) sigma[i * NUMBER_OF_IONNAMES + ion] = ph.get_cross_section(ion);
    she[i] = ph.get_cross_section_He_corr();
    weight[i] = ph.get_weight();
  }
}

/* n x PhotonSource::reemit (src/PhotonSource.cpp:272-308) with RandomGenerator(seed): packet i was absorbed at
 * frequency nu_in[i] in a cell with (xH, xHe, T)[i] */
void cmi_ref_reemit_sequence(const char *paramfile, int seed, int64_t n, const double *xH, const double *xHe,
                             const double *T, const double *nu_in, double *nu_out, int32_t *type_out, double *dir) {
  RefPhotonSource s(paramfile);
  RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i) {
    Photon ph(CoordinateVector<>(0.), CoordinateVector<>(1., 0., 0.), nu_in[i]);
    s.source->set_cross_sections(ph, nu_in[i]);
    IonizationVariables iv;
    iv.set_temperature(T[i]);
    iv.set_ionic_fraction(ION_H_n, xH[i]);
    iv.set_ionic_fraction(ION_He_n, xHe[i]);
    s.source->_reemission_handler->set_reemission_probabilities(iv);
    const bool ok = s.source->reemit(ph, iv, rg);
    nu_out[i] = ok ? ph.get_energy() : 0.;
    type_out[i] = (int32_t)ph.get_type();
    for (int k = 0; k < 3; ++k) dir[3 * i + k] = ok ? ph.get_direction()[k] : 0.;
  }
}

/* FaucherGiguerePhotonSourceSpectrum(redshift) (src/FaucherGiguerePhotonSourceSpectrum.cpp): its
 * frequency grid and cumulative distribution (what a tabulated spectrum hands to the device),
 * its total flux, and n samples with RandomGenerator(seed) together with the deviates they
 * consumed (one per sample).  which = 1: UniformPhotonSourceSpectrum instead (no tables). */
int cmi_ref_tabulated_spectrum(int which, double redshift, int seed, int64_t n, double *uniforms, double *nu,
                               double *freq, double *cdf, int capacity, double *total_flux) {
  RandomGenerator rg(seed), replay(seed);
  for (int64_t i = 0; i < n; ++i) uniforms[i] = replay.get_uniform_random_double();
  if (which == 1) {
    UniformPhotonSourceSpectrum sp;
    for (int64_t i = 0; i < n; ++i) nu[i] = sp.get_random_frequency(rg, 0.);
    return 0;
  }
  FaucherGiguerePhotonSourceSpectrum sp(redshift);
  const int m = (int)sp._frequencies.size();
  if (m > capacity) return -m;
  for (int i = 0; i < m; ++i) {
    freq[i] = sp._frequencies[i];
    cdf[i] = sp._cumulative_distribution[i];
  }
  if (total_flux) *total_flux = sp.get_total_flux();
  for (int64_t i = 0; i < n; ++i) nu[i] = sp.get_random_frequency(rg, 0.);
  return m;
}

/* RandomGenerator(seed) (src/RandomGenerator.hpp): n deviates */
void cmi_ref_random_stream(int seed, int64_t n, double *out) {
  RandomGenerator rg(seed);
  for (int64_t i = 0; i < n; ++i) out[i] = rg.get_uniform_random_double();
}

/* PhotonSourceDistributionFactory::generate on a parameter file: info = {number of sources, total
 * luminosity}; positions[capacity][3], weights[capacity].  Returns the number of sources, -1 for None. */
int cmi_ref_photon_source_distribution(const char *paramfile, double *info, double *positions, double *weights,
                                       int capacity) {
  ParameterFile params(paramfile);
  PhotonSourceDistribution *d = PhotonSourceDistributionFactory::generate(params, nullptr);
  if (d == nullptr) return -1;
  const int n = (int)d->get_number_of_sources();
  info[0] = n;
  info[1] = d->get_total_luminosity();
  for (int i = 0; i < n && i < capacity; ++i) {
    const CoordinateVector<> x = d->get_position(i);
    positions[3 * i] = x.x(); positions[3 * i + 1] = x.y(); positions[3 * i + 2] = x.z();
    weights[i] = d->get_weight(i);
  }
  delete d;
  return n;
}

/* SPHArrayInterface (src/SPHArrayInterface.cpp) on the parameter file's grid: the densities
 * IonizationSimulation::initialize(interface) maps onto the cells, and the inverse mapping
 * (write + fill_array) of a given neutral-fraction field.  box_anchor == NULL: non-periodic interface. */
int64_t cmi_ref_sph_mapping(const char *paramfile, const char *mapping_type, const double *box_anchor,
                            const double *box_sides, int64_t N, const double *x, const double *y, const double *z,
                            const double *h, const double *m, int64_t ncell, double *dens, const double *xH_cells,
                            double *nH) {
  SPHArrayInterface *sph = box_anchor ? new SPHArrayInterface(1., 1., box_anchor, box_sides, mapping_type)
                                      : new SPHArrayInterface(1., 1., mapping_type);
  IonizationSimulation sim(false, false, false, 1, paramfile, nullptr, nullptr);
  sph->reset(x, y, z, h, m, (size_t)N);
  sim.initialize(sph);
  DensityGrid &grid = *sim._density_grid;
  const int64_t n = grid.get_number_of_cells();
  if (n == ncell) {
    int64_t i = 0;
    for (auto it = grid.begin(); it != grid.end(); ++it, ++i) {
      dens[i] = it.get_ionization_variables().get_number_density();
      it.get_ionization_variables().set_ionic_fraction(ION_H_n, xH_cells[i]);
    }
    sph->write(grid, 0, sim._parameter_file, 0., nullptr);
    sph->fill_array(nH);
  }
  delete sph;
  return n;
}

/* The SPH kernel sum of GadgetSnapshotDensityFunction::operator() (GadgetSnapshotDensityFunction.cpp:315-359)
 * on particle arrays handed in directly (the class itself only reads HDF5 files, and this build has no HDF5):
 * the reference's Octree (built as the constructor does, :226-251, periodic or in the 1 % padded bounding box)
 * finds the particles whose kernel contains a point, the reference's CubicSplineKernel weighs them.
 * out[3q..3q+3) = number density, temperature, neutral fraction (or -1 without neutral fractions). */
void cmi_ref_gadget_kernel_sums(int64_t N, const double *pos, const double *m, const double *h, const double *rho,
                                const double *T, const double *xH, int periodic, const double *box_sides, int64_t nq,
                                const double *q, double *out) {
  std::vector< CoordinateVector<> > positions(N);
  std::vector< double > hs(h, h + N);
  CoordinateVector<> minpos(DBL_MAX), maxpos(-DBL_MAX);
  for (int64_t i = 0; i < N; ++i) {
    positions[i] = CoordinateVector<>(pos[3 * i], pos[3 * i + 1], pos[3 * i + 2]);
    minpos = CoordinateVector<>::min(minpos, positions[i]);
    maxpos = CoordinateVector<>::max(maxpos, positions[i]);
  }
  Box<> pbox;
  if (periodic) pbox = Box<>(CoordinateVector<>(), CoordinateVector<>(box_sides[0], box_sides[1], box_sides[2]));
  Box<> box(pbox);
  if (!periodic) {
    CoordinateVector<> sides = maxpos - minpos;
    const CoordinateVector<> anchor = minpos - 0.005 * sides;
    sides *= 1.01;
    box = Box<>(anchor, sides);
  }
  Octree octree(positions, box, periodic != 0);
  octree.set_auxiliaries(hs, Octree::max< double >);
  for (int64_t k = 0; k < nq; ++k) {
    const CoordinateVector<> position(q[3 * k], q[3 * k + 1], q[3 * k + 2]);
    double density = 0., temperature = 0., neutral_fraction = xH ? 0. : -1.;
    const std::vector< uint_fast32_t > ngbs = octree.get_ngbs(position);
    for (size_t i = 0; i < ngbs.size(); ++i) {
      const uint_fast32_t index = ngbs[i];
      double r;
      if (!pbox.get_sides().x()) r = (position - positions[index]).norm();
      else r = pbox.periodic_distance(position, positions[index]).norm();
      const double u = r / hs[index];
      const double splineval = m[index] * CubicSplineKernel::kernel_evaluate(u, hs[index]);
      density += splineval;
      temperature += splineval * T[index] / rho[index];
      if (neutral_fraction >= 0.) neutral_fraction += splineval * xH[index];
    }
    out[3 * k] = density / 1.6737236e-27;
    out[3 * k + 1] = temperature;
    out[3 * k + 2] = (neutral_fraction >= 0.) ? neutral_fraction / density : -1.;
  }
}

/* UnitConverter::to_SI for a quantity given by its SI unit name, e.g.
 * ("13.6", "eV" -> "Hz").  Used to pin the host-side unit parser. */
double cmi_ref_convert(double value, const char *unit_from, const char *unit_to) {
  return UnitConverter::convert(value, unit_from, unit_to);
}


/* The reference's ParameterFile on a file: query `nkeys` keys (newline separated list
 * "kind|key|default", kind in s(tring) d(ouble) b(ool) i(nteger) and the physical quantities
 * L(ength) N(umber density) T(emperature) F(requency) A(rea) R(eaction rate)), write the values
 * as text lines into `out`, followed by the used-values dump (print_contents).  Pins the
 * host-side parameter-file parser of the product. */
int cmi_ref_paramfile_query(const char *filename, const char *queries, char *out, int nout) {
  ParameterFile params(filename);
  std::stringstream in(queries), res;
  res.precision(17);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    const char kind = line[0];
    const size_t bar = line.find('|', 2);
    const std::string key = line.substr(2, bar - 2);
    const std::string def = line.substr(bar + 1);
    switch (kind) {
    case 's': res << params.get_value< std::string >(key, def); break;
    case 'd': res << params.get_value< double >(key, atof(def.c_str())); break;
    case 'b': res << params.get_value< bool >(key, def == "true"); break;
    case 'i': res << params.get_value< uint_fast64_t >(key, (uint_fast64_t)atof(def.c_str())); break;
    case 'L': res << params.get_physical_value< QUANTITY_LENGTH >(key, def); break;
    case 'N': res << params.get_physical_value< QUANTITY_NUMBER_DENSITY >(key, def); break;
    case 'T': res << params.get_physical_value< QUANTITY_TEMPERATURE >(key, def); break;
    case 'F': res << params.get_physical_value< QUANTITY_FREQUENCY >(key, def); break;
    case 'A': res << params.get_physical_value< QUANTITY_SURFACE_AREA >(key, def); break;
    case 'R': res << params.get_physical_value< QUANTITY_REACTION_RATE >(key, def); break;
    default: res << "?";
    }
    res << "\n";
  }
  res << "---\n";
  std::stringstream dump;
  params.print_contents(dump);
  /* drop the time-stamp line */
  std::string d = dump.str();
  d = d.substr(d.find('\n') + 1);
  res << d;
  const std::string r = res.str();
  strncpy(out, r.c_str(), nout - 1);
  out[nout - 1] = 0;
  return (int)r.size();
}


/* ---------------------------------------------------------------------------
 * Step-by-step driving of the reference IonizationSimulation: the body of the
 * `while (loop < _number_of_iterations)` loop of IonizationSimulation::run
 * (IonizationSimulation.cpp:359-643, MPI branches excluded) executed one
 * iteration per call ON THE REFERENCE'S OWN OBJECTS, so that each phase can be
 * timed separately and the grid state can be read or replaced in between.
 *   reset_grid                                   :379
 *   set_reemission_probabilities(grid)           :380-383
 *   set_numphoton + do_in_parallel(shoot market) :399-404
 *   update_counters                              :406
 *   calculate_temperature(loop, totweight, ...)  :532-533
 * out[0] = shoot seconds, out[1] = state-update seconds, out[2] = totweight,
 * out[3..6] = typecount, out[7] = reset + reemission-probability seconds.
 * ------------------------------------------------------------------------- */
struct cmi_ref_sim {
  IonizationSimulation *sim;
  std::pair<cellsize_t, cellsize_t> block;
};

void *cmi_ref_sim_create(const char *paramfile, int num_threads) {
  cmi_ref_sim *h = new cmi_ref_sim();
  h->sim = new IonizationSimulation(false, false, false, num_threads, paramfile, nullptr, nullptr);
  h->sim->initialize();
  h->block = std::make_pair((cellsize_t)0, h->sim->_density_grid->get_number_of_cells());
  return h;
}

void cmi_ref_sim_destroy(void *handle) {
  cmi_ref_sim *h = (cmi_ref_sim *)handle;
  delete h->sim;
  delete h;
}

int64_t cmi_ref_sim_number_of_cells(void *handle) {
  return ((cmi_ref_sim *)handle)->sim->_density_grid->get_number_of_cells();
}

int cmi_ref_sim_threads(void *handle) { return ((cmi_ref_sim *)handle)->sim->_num_thread; }

int cmi_ref_sim_iteration(void *handle, uint32_t loop, uint64_t numphoton, double *out) {
  IonizationSimulation &s = *((cmi_ref_sim *)handle)->sim;
  Timer t_prep, t_shoot, t_update;
  t_prep.start();
  s._density_grid->reset_grid(*s._density_function);
  if (s._photon_source->get_reemission_handler())
    s._photon_source->get_reemission_handler()->set_reemission_probabilities(*s._density_grid);
  t_prep.stop();
  double typecount[PHOTONTYPE_NUMBER] = {0};
  double totweight = 0.;
  s._ionization_photon_shoot_job_market->set_numphoton(numphoton);
  t_shoot.start();
  s._work_distributor.do_in_parallel(*s._ionization_photon_shoot_job_market);
  t_shoot.stop();
  s._ionization_photon_shoot_job_market->update_counters(totweight, typecount);
  t_update.start();
  s._temperature_calculator->calculate_temperature(loop, totweight, *s._density_grid,
                                                   ((cmi_ref_sim *)handle)->block);
  t_update.stop();
  out[0] = t_shoot.value();
  out[1] = t_update.value();
  out[2] = totweight;
  for (int k = 0; k < PHOTONTYPE_NUMBER; ++k) out[3 + k] = typecount[k];
  out[7] = t_prep.value();
  return 0;
}

/* fields [32][ncell]: n, T, x[14], raw J[14] (of the last shoot), heat[2] */
void cmi_ref_sim_get_fields(void *handle, double *fields) {
  DensityGrid &grid = *((cmi_ref_sim *)handle)->sim->_density_grid;
  const int64_t n = grid.get_number_of_cells();
  for (auto it = grid.begin(); it != grid.end(); ++it) {
    const int64_t i = it.get_index();
    const IonizationVariables &iv = it.get_ionization_variables();
    fields[0 * n + i] = iv.get_number_density();
    fields[1 * n + i] = iv.get_temperature();
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion) {
      fields[(2 + ion) * n + i] = iv.get_ionic_fraction(ion);
      fields[(16 + ion) * n + i] = iv.get_mean_intensity(ion);
    }
    fields[30 * n + i] = iv.get_heating(HEATINGTERM_H);
    fields[31 * n + i] = iv.get_heating(HEATINGTERM_He);
  }
}

/* overwrite n, T, x[14] of every cell (fields [16][ncell]) */
void cmi_ref_sim_set_state(void *handle, const double *fields) {
  DensityGrid &grid = *((cmi_ref_sim *)handle)->sim->_density_grid;
  const int64_t n = grid.get_number_of_cells();
  for (auto it = grid.begin(); it != grid.end(); ++it) {
    const int64_t i = it.get_index();
    IonizationVariables &iv = it.get_ionization_variables();
    iv.set_number_density(fields[0 * n + i]);
    iv.set_temperature(fields[1 * n + i]);
    for (int ion = 0; ion < NUMBER_OF_IONNAMES; ++ion)
      iv.set_ionic_fraction(ion, fields[(2 + ion) * n + i]);
  }
}

} /* extern "C" */
