/*
 * cmib.h — C ABI of the B200-native photoionization hot path ("CMacIonize
 * B200 backend").  This is the drop-in boundary: plain pointers and sizes, no
 * C++ or torch types.  Every entry point names the reference interface it
 * replaces (paths relative to the reference tree, bwvdnbro/CMacIonize).
 *
 * The reference has no fine-grained FFI for this path; the path sits behind C++
 * abstract classes chosen by parameter-file `type:` strings and is driven by
 * IonizationSimulation::run (src/IonizationSimulation.cpp:334-679).  The entry
 * points below are what a maintainer would call from the bodies of those
 * classes (INTEGRATION.md shows the shim for each).
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on error; cmib_last_error()
 *    returns the message.  The reference aborts on every error
 *    (src/Error.hpp:101-106 `cmac_error`); set CMIB_ABORT_ON_ERROR=1 in the
 *    environment or call cmib_set_abort_on_error(1) for the same behaviour.
 *  - all physical quantities are SI, exactly as inside the reference.
 *  - host arrays belong to the caller, device memory to the library.
 *  - per-cell arrays use the reference's cell order: long index
 *    ix*ny*nz + iy*nz + iz (src/CartesianDensityGrid.hpp:137-144); per-ion
 *    arrays are [ion][cell] with the reference's IonName order
 *    (src/ElementNames.hpp:107-160): H0 He0 C+ C++ N0 N+ N++ O0 O+ Ne0 Ne+ S+
 *    S++ S+++.
 *  - one host thread per context; a context owns one CUDA stream on one device.
 *  - there is NO CPU fallback: every call fails if no sm_100 device is present.
 */
#ifndef CMIB_H
#define CMIB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMIB_ABI_VERSION 1
#define CMIB_NUM_IONS 14
#define CMIB_NUM_HEATING_TERMS 2
#define CMIB_NUM_ELEMENTS 6 /* He C N O Ne S (src/ElementNames.hpp:52-88) */
#define CMIB_NUM_REEMISSION_PROBABILITIES 5
#define CMIB_NUM_PACKET_TYPES 4 /* src/PhotonType.hpp:41-56 */

typedef struct cmib_context cmib_context;

/* SimulationBox + CartesianDensityGrid constructor arguments
 * (src/SimulationBox.hpp:63-72, src/CartesianDensityGrid.cpp:44-92) */
typedef struct cmib_grid_desc {
  double anchor[3];    /* SimulationBox:anchor (m) */
  double sides[3];     /* SimulationBox:sides (m) */
  int32_t ncell[3];    /* DensityGrid:number of cells */
  int32_t periodic[3]; /* SimulationBox:periodicity */
} cmib_grid_desc;

/* plugin selectors (the reference's `type:` strings) */
enum { CMIB_CROSS_SECTIONS_FIXED_VALUE = 0, CMIB_CROSS_SECTIONS_VERNER = 1 };
enum { CMIB_RECOMBINATION_FIXED_VALUE = 0, CMIB_RECOMBINATION_VERNER = 1 };
enum { CMIB_SPECTRUM_MONOCHROMATIC = 0, CMIB_SPECTRUM_PLANCK = 1, CMIB_SPECTRUM_UNIFORM = 2, CMIB_SPECTRUM_TABULATED = 3 };
enum { CMIB_CONTINUOUS_NONE = 0, CMIB_CONTINUOUS_ISOTROPIC = 1, CMIB_CONTINUOUS_PLANAR = 2, CMIB_CONTINUOUS_DISTANT_STAR = 3,
       CMIB_CONTINUOUS_EXTENDED_DISC = 4, CMIB_CONTINUOUS_SPIRAL_GALAXY = 5 };
enum { CMIB_REEMISSION_NONE = 0, CMIB_REEMISSION_PHYSICAL = 1, CMIB_REEMISSION_FIXED_VALUE = 2 };

/* TemperatureCalculator parameters (src/TemperatureCalculator.cpp:133-160) */
typedef struct cmib_temperature_params {
  int32_t do_temperature_calculation;
  uint32_t minimum_number_of_iterations;
  double epsilon_convergence;
  uint32_t maximum_number_of_iterations;
  double pah_heating_factor;
  double cosmic_ray_heating_factor;
  double cosmic_ray_heating_limit;
  double cosmic_ray_heating_scale_length; /* m */
  double minimum_ionized_temperature;     /* K */
} cmib_temperature_params;

/* ---- library / context ------------------------------------------------- */
int cmib_abi_version(void);
const char *cmib_last_error(void);
void cmib_set_abort_on_error(int on);
/* number of kernels this library has launched in this process (bench evidence) */
uint64_t cmib_kernel_launch_count(void);

/* replaces: CartesianDensityGrid::CartesianDensityGrid + DensityGrid::allocate_memory
 * (src/CartesianDensityGrid.cpp:44-92, src/DensityGrid.hpp:235-259) */
int cmib_create(const cmib_grid_desc *grid, int device, cmib_context **out);
int cmib_destroy(cmib_context *ctx);
int cmib_synchronize(cmib_context *ctx);

/* ---- grid state -------------------------------------------------------- */
/* replaces: DensityGrid::set_densities (src/DensityGrid.cpp:40-62, per-cell functor
 * src/DensityGrid.hpp:775-790).  x is [14][ncell]; cosmic_ray_factor may be NULL
 * (reference default -1, src/IonizationVariables.hpp:125). */
int cmib_upload_cells(cmib_context *ctx, const double *number_density, const double *temperature,
                      const double *ionic_fractions, const double *cosmic_ray_factor);
/* replaces: the IonizationVariables getters a DensityGridWriter walks
 * (src/IonizationVariables.hpp:226-345).  Any pointer may be NULL.  heating is
 * [2][ncell], the normalised heating terms the state update leaves in the cell. */
int cmib_download_cells(cmib_context *ctx, double *number_density, double *temperature,
                        double *ionic_fractions, double *heating);
/* raw mean-intensity / heating sums of the current iteration, J [14][ncell],
 * heat [2][ncell] (IonizationVariables::get_mean_intensity / get_heating before
 * normalisation).  Either may be NULL. */
int cmib_download_accumulators(cmib_context *ctx, double *mean_intensity, double *heating);
/* replaces: DensityGrid::reset_grid (src/DensityGrid.hpp:803-807) */
int cmib_reset_accumulators(cmib_context *ctx);

/* ---- plugins ----------------------------------------------------------- */
/* Abundances (src/Abundances.hpp:53-76): He C N O Ne S relative to H */
int cmib_set_abundances(cmib_context *ctx, const double abundances[CMIB_NUM_ELEMENTS]);
/* The reference has two drivers with different conventions for what a packet carries (same physics):
 *   CMIB_CONVENTIONS_IONIZATION_SIMULATION (default): plain cross sections; heating terms relative to 13.6 eV and
 *     24.6 eV converted to Hz (src/DensityGrid.hpp:219-222);
 *   CMIB_CONVENTIONS_TASK_BASED (`CMacIonize --task-based`): the abundance of an ion's element is folded into the cross
 *     section of every ion but H0 (src/SourceDiscretePhotonTaskContext.hpp:172-180), the mean intensities are divided by
 *     it again before the state update, the helium heating term by the helium abundance
 *     (src/TaskBasedIonizationSimulation.cpp:932-951), the re-emission decision uses A_He = 1
 *     (src/PhotonReemitTaskContext.hpp:121-127), and the heating terms use the hard-coded thresholds 3.288e15 Hz and
 *     5.948e15 Hz (src/DensitySubGrid.hpp:608-612).
 * cmib_download_accumulators returns what the shoot accumulated (abundance-weighted under the second convention). */
#define CMIB_CONVENTIONS_IONIZATION_SIMULATION 0
#define CMIB_CONVENTIONS_TASK_BASED 1
int cmib_set_packet_conventions(cmib_context *ctx, int conventions);
/* CrossSectionsFactory (src/CrossSectionsFactory.hpp:60-80); fixed[14] in m^2 is read
 * for FIXED_VALUE (src/FixedValueCrossSections.hpp), ignored for VERNER */
int cmib_set_cross_sections(cmib_context *ctx, int kind, const double fixed[CMIB_NUM_IONS]);
/* CrossSections:type Bimodal (src/BimodalCrossSections.hpp:247-254): low[ion] below the frequency limit (Hz),
 * high[ion] at and above it, both in m^2 */
int cmib_set_bimodal_cross_sections(cmib_context *ctx, double frequency_limit, const double low[CMIB_NUM_IONS],
                                    const double high[CMIB_NUM_IONS]);
/* RecombinationRatesFactory (src/RecombinationRatesFactory.hpp:59-72); fixed[14] m^3 s^-1 */
int cmib_set_recombination_rates(cmib_context *ctx, int kind, const double fixed[CMIB_NUM_IONS]);
/* PhotonSourceDistribution -> PhotonSource (src/PhotonSource.cpp:55-146): positions
 * [n][3] (m), weights [n] summing to 1 (checked to 1e-9 like the reference),
 * total luminosity (s^-1) */
int cmib_set_sources(cmib_context *ctx, int32_t n_sources, const double *positions,
                     const double *weights, double total_luminosity);
/* PhotonSourceSpectrumFactory (src/PhotonSourceSpectrumFactory.hpp:84-152): param is the
 * frequency (Hz) for MONOCHROMATIC, the black-body temperature (K) for PLANCK, unused for UNIFORM
 * (src/UniformPhotonSourceSpectrum.hpp:50-53: 13.6 - 54.4 eV) */
int cmib_set_spectrum(cmib_context *ctx, int kind, double param);
/* Any spectrum that samples from a frequency grid and its cumulative distribution with linear
 * interpolation — FaucherGiguere (src/FaucherGiguerePhotonSourceSpectrum.cpp:234-247), WMBasic
 * (src/WMBasicPhotonSourceSpectrum.cpp:236-248), PopStar, Pegase3, CastelliKurucz, Masked
 * (src/MaskedPhotonSourceSpectrum.cpp:123-135): hand over its two arrays (frequencies in Hz, cumulative
 * distribution from 0 to 1, both [n]).  role 0 = PhotonSourceSpectrum of the discrete sources,
 * 1 = ContinuousPhotonSourceSpectrum (then pass CMIB_SPECTRUM_TABULATED to cmib_set_continuous_source). */
int cmib_set_spectrum_table(cmib_context *ctx, int role, int32_t n, const double *frequencies,
                            const double *cumulative_distribution);
/* ContinuousPhotonSourceFactory (src/ContinuousPhotonSourceFactory.hpp:69-100) + its spectrum
 * (PhotonSourceSpectrumFactory with role "ContinuousPhotonSourceSpectrum"): kind ISOTROPIC =
 * IsotropicContinuousPhotonSource (src/IsotropicContinuousPhotonSource.hpp:106-180: packets enter
 * through the faces of the box); luminosity (s^-1) = total surface area x total flux of the spectrum
 * (PhotonSource.cpp:101-108); spectrum_kind / spectrum_param as for cmib_set_spectrum.  With
 * discrete sources present half of the packets come from the continuous source and carry the
 * weight L_continuous / L_discrete (PhotonSource.cpp:113-131); cmib_set_sources may be called with
 * n_sources = 0 (PhotonSourceDistribution: None) when this is set.  kind NONE removes it. */
int cmib_set_continuous_source(cmib_context *ctx, int kind, double luminosity, int spectrum_kind,
                               double spectrum_param);
/* geometry of a PLANAR continuous source (src/PlanarContinuousPhotonSource.hpp:101-131): an
 * isotropically emitting rectangle in the plane coordinate[normal_axis] = intercept, spanning
 * anchor[k] .. anchor[k] + sides[k] in the two other coordinates (in ascending coordinate order);
 * call before cmib_set_continuous_source(ctx, CMIB_CONTINUOUS_PLANAR, luminosity, ...), whose luminosity
 * is the source's own (ContinuousPhotonSource:luminosity), not area x flux */
/* position (m) of a DISTANT_STAR continuous source (src/DistantStarContinuousPhotonSource.hpp:60-90): a star
 * outside the box that illuminates the faces turned towards it; it must lie outside the box.  Call before
 * cmib_set_continuous_source(ctx, CMIB_CONTINUOUS_DISTANT_STAR, luminosity, ...) with luminosity = exposed
 * surface area x total flux (PhotonSource.cpp:104-108, get_total_surface_area :194-212) */
int cmib_set_distant_star_position(cmib_context *ctx, const double position[3]);
/* ExtendedDiscContinuousPhotonSource(box, params) (src/ExtendedDiscContinuousPhotonSource.hpp:102-117): emission from
 * the volume of a disc, Gaussian with `scale_height` around coordinate[normal_axis] = origin, uniform over the box in
 * the two other coordinates; call before cmib_set_continuous_source(ctx, CMIB_CONTINUOUS_EXTENDED_DISC, luminosity, ...),
 * whose luminosity is the source's own `luminosity` key (has_total_luminosity(), PhotonSource.cpp:101-103). */
int cmib_set_extended_disc_geometry(cmib_context *ctx, int normal_axis, double origin, double scale_height);
/* SpiralGalaxyContinuousPhotonSource(box, params) (src/SpiralGalaxyContinuousPhotonSource.hpp:78-120): stellar bulge +
 * double exponential disc around the origin (scale length and height of the disc, bulge over total ratio); the radial
 * luminosity table is built here.  Call before cmib_set_continuous_source(ctx, CMIB_CONTINUOUS_SPIRAL_GALAXY,
 * luminosity, ...) with luminosity = get_total_surface_area() (= 1 m^2, :194) x total flux of the spectrum. */
int cmib_set_spiral_galaxy_geometry(cmib_context *ctx, double scale_length_stars, double scale_height_stars,
                                    double bulge_over_total_ratio);
int cmib_set_planar_source_geometry(cmib_context *ctx, int normal_axis, double intercept, const double anchor[2],
                                    const double sides[2]);
/* DiffuseReemissionHandlerFactory (src/DiffuseReemissionHandlerFactory.hpp:59-107);
 * probability / frequency (Hz) are used by FIXED_VALUE only.  Builds the H-Lyc /
 * He-Lyc / He-2-photon tables from the CURRENT cross sections, as the reference's
 * PhysicalDiffuseReemissionHandler constructor does: set cross sections first. */
int cmib_set_reemission(cmib_context *ctx, int kind, double probability, double frequency);
int cmib_set_temperature_params(cmib_context *ctx, const cmib_temperature_params *params);

/* ---- one photoionization iteration (src/IonizationSimulation.cpp:359-643) ---- */
/* replaces: DiffuseReemissionHandler::set_reemission_probabilities(grid)
 * (src/PhysicalDiffuseReemissionHandler.hpp:66-106; call site IonizationSimulation.cpp:380-383) */
int cmib_update_reemission_probabilities(cmib_context *ctx);
/* replaces: WorkDistributor::do_in_parallel(IonizationPhotonShootJobMarket)
 * (src/IonizationSimulation.cpp:399-406; per packet src/IonizationPhotonShootJob.hpp:117-146).
 * Shoots packets with global ids [packet_offset, packet_offset + n_packets); the
 * RNG stream of a packet depends only on (seed, iteration, global id), so
 * splitting a batch over GPUs does not change which packets are drawn.
 * totweight / typecount[4] (may be NULL: no host synchronisation then) receive
 * this call's sums (IonizationPhotonShootJobMarket::update_counters). */
int cmib_shoot(cmib_context *ctx, uint64_t n_packets, uint64_t packet_offset, uint64_t seed,
               uint32_t iteration, double *totweight, double *typecount);
/* replaces: TemperatureCalculator::calculate_temperature(loop, totweight, grid, block)
 * (src/TemperatureCalculator.cpp:944-970), i.e. the temperature solve when enabled
 * and loop > minimum number of iterations, else
 * IonizationStateCalculator::calculate_ionization_state (src/IonizationStateCalculator.cpp:511-530).
 * totweight <= 0 means "use the device-side sum" (after an all-reduce of the
 * accumulator buffer, so no host round trip is needed). */
int cmib_update_state(cmib_context *ctx, uint32_t loop, double totweight);

/* diagnostics of the shoots since the last cmib_reset_accumulators: number of
 * packet-cell crossings (the unit of work of the roofline) and of
 * (re-)emissions.  No reference counterpart (gprof call counts were used there). */
int cmib_shoot_statistics(cmib_context *ctx, double *cell_crossings, double *emissions);
/* optical depth traversed by all packets since the last reset: sum over crossings of
 * ds n (sigma_H x_H + A_He sigma_He x_He) (the shortened last crossing counts what was left).
 * Checksum for the accumulation at any problem size: it must equal
 * sum_cells n (x_H J_H + A_He x_He J_He) of the raw accumulators (unit packet weights). */
int cmib_shoot_optical_depth(cmib_context *ctx, double *tau_traversed);

/* per-kernel device times of the LAST cmib_shoot (CUDA events on the context's stream; enable with
 * cmib_set_shoot_timing before the shoot): summed prepare / march kernel milliseconds, number of
 * rounds, and the number of accumulator terms added since the last reset (RED operations, the
 * unit of the atomic roofline).  Any pointer may be NULL.  No reference counterpart. */
int cmib_set_shoot_timing(cmib_context *ctx, int on);
int cmib_shoot_timing(cmib_context *ctx, double *prepare_ms, double *march_ms, uint64_t *rounds,
                      double *accumulator_adds);
/* A large shoot runs as two lanes (two sets of queues on two streams, out of phase) so that the emission kernels
 * of one lane execute beside the march kernel of the other.  With lanes the times of cmib_shoot_timing are the
 * durations during which AT LEAST ONE emission kernel (prepare_ms) / march kernel (march_ms) was running (unions of
 * the kernels' intervals on a common clock); this call adds the number of lanes of the last shoot and the time
 * during which both kinds ran at once (prepare_ms + march_ms - overlap_ms = time with any of them running). */
int cmib_shoot_overlap(cmib_context *ctx, int32_t *lanes, double *overlap_ms);

/* test hook: 0 = wavefront pipeline (default, production: prepare/march kernels connected by
 * device queues), 1 = one-thread-per-packet kernel.  Both draw the same packets from the same
 * per-packet random streams; they differ only in the order of the atomic adds. */
int cmib_set_shoot_algorithm(cmib_context *ctx, int algorithm);

/* ---- multi-GPU plumbing ------------------------------------------------ */
/* Device pointer + length (in doubles) of the contiguous buffer that must be
 * sum-all-reduced between cmib_shoot and cmib_update_state: 16 counters
 * (totweight, typecount[4], cell crossings, emissions, accumulator adds, optical depth
 * traversed, padding to one 128-byte line) followed by the per-cell accumulators.
 * Replaces the 16 chunked MPI_Allreduce calls + 2 counter reductions of
 * src/IonizationSimulation.cpp:410-416,458-529 with ONE collective. */
int cmib_accumulator_buffer(cmib_context *ctx, void **device_ptr, uint64_t *n_doubles);
/* raw cudaStream_t of the context (for ordering an external collective) */
int cmib_stream(cmib_context *ctx, void **stream);

/* ---- multi-GPU: the reference's MPICommunicator for this path ----------- */
/* The reference splits an iteration over MPI ranks (src/IonizationSimulation.cpp:392-397, 458-618): every rank
 * shoots distribute(numphoton) packets on a replicated grid, 16 chunked MPI_Allreduce calls sum the per-cell
 * accumulators, every rank updates the cell block distribute_block gives it, and 15 hand-rolled all-gathers
 * rebuild the replicas.  Same three steps here, with a load-balanced deal of the cells (below).  Here one context = one GPU = one rank; the collectives are NCCL calls on the context's
 * stream (libnccl.so.2 is bound at run time: single-GPU users need no NCCL).
 *
 * cmib_distribute / cmib_distribute_block: MPICommunicator::distribute (MPICommunicator.hpp:207-222) and
 * ::distribute_block (:237-255), pure functions (no context, no GPU). */
uint64_t cmib_distribute(uint64_t number, int32_t size, int32_t rank);
void cmib_distribute_block(int32_t rank, int32_t size, uint64_t begin, uint64_t end, uint64_t *block_begin,
                           uint64_t *block_end);
/* One communicator over `size` contexts.  Several processes (one rank each, e.g. under torchrun or mpirun):
 * rank 0 calls cmib_comm_unique_id and hands the 128 bytes to the others by any means, then every rank calls
 * cmib_comm_init_rank (collective: it blocks until all ranks have called it).  One process driving several
 * GPUs from one host thread each: cmib_comm_init_all on the array of contexts. */
int cmib_comm_unique_id(void *id128);
int cmib_comm_init_rank(cmib_context *ctx, int32_t size, int32_t rank, const void *id128);
int cmib_comm_init_all(cmib_context **ctxs, int32_t size);
int cmib_comm_finalize(cmib_context *ctx);
/* rank, size and the cell block [begin, end) of cmib_distribute_block for this rank: the block a distributed caller
 * uploads (cmib_upload_cells_block) — size 1 and the whole grid without a communicator */
int cmib_comm_info(cmib_context *ctx, int32_t *rank, int32_t *size, uint64_t *cell_begin, uint64_t *cell_end);
/* The exchange between cmib_shoot and the next iteration, as ONE call per rank (collective):
 *   1. the accumulators (and the 16 leading counters) are summed over the ranks: one ncclAllReduce, every rank
 *      receives all sums (`allreduce` is kept for source compatibility and ignored); the heating plane of the H-only
 *      layout is left out when nothing shot since cmib_reset_accumulators could add to it;
 *   2. cmib_update_state on the cells the rank OWNS (totweight from the reduced counters).  Ownership is dealt in
 *      chunks of 1024 cells, chunk c to rank c % size (cmib_owned_cell): the reference's contiguous blocks
 *      (distribute_block) put the whole ionised region — where the temperature solve iterates — on the middle
 *      ranks (measured on lexingtonHII20, 8 GPUs: 2.1 ms for the slowest block against 0.6 ms for an eighth of
 *      the work); the result per cell does not depend on who computes it;
 *   3. the opacity records (n, x_H, x_He, T) of all chunks are gathered on every rank (equal-sized packs of the
 *      owned chunks, one ncclAllGather): all a shoot reads.
 * The remaining per-cell state (metal fractions, heating terms) stays distributed — ownership never changes —
 * until cmib_comm_gather_state is called (before a rank downloads cells it does not own).
 * Without a communicator the call is cmib_update_state on the whole grid. */
int cmib_comm_exchange_and_update(cmib_context *ctx, uint32_t loop, int allreduce);
int cmib_comm_gather_state(cmib_context *ctx);
/* milliseconds (CUDA events on the context's stream) of the three phases of the last exchange: reduce, update, gather */
int cmib_comm_exchange_timing(cmib_context *ctx, double ms[3]);
/* cmib_update_state restricted to the cells [cell_begin, cell_end) */
int cmib_update_state_block(cmib_context *ctx, uint32_t loop, double totweight, uint64_t cell_begin, uint64_t cell_end);
/* upload / download of a cell block from / to host arrays that hold ONLY that block (same field layout as
 * cmib_upload_cells / cmib_download_cells with ncell = cell_end - cell_begin): the distributed form of the
 * end-to-end path — every rank moves 1/size of the bytes, cmib_comm_gather_cells_all replicates what was uploaded */
int cmib_upload_cells_block(cmib_context *ctx, uint64_t cell_begin, uint64_t cell_end, const double *number_density,
                            const double *temperature, const double *ionic_fractions);
int cmib_download_cells_block(cmib_context *ctx, uint64_t cell_begin, uint64_t cell_end, double *number_density,
                              double *temperature, double *ionic_fractions, double *heating);
/* all-gather of everything cmib_upload_cells_block wrote (opacity records + metal fractions); the blocks are those
 * of cmib_distribute_block over the communicator */
int cmib_comm_gather_cells_all(cmib_context *ctx);
/* the cells a rank owns in the state update: work item j of rank `rank` of `size` is cell cmib_owned_cell(j, size, rank)
 * (pure functions); cmib_comm_owned_cells = cmib_owned_cell_count of the context's grid and rank;
 * cmib_download_cells_owned reads the owned cells back in work-item order (field layout of cmib_download_cells with
 * ncell = the number of owned cells): the distributed read-back of an iteration, no gather needed */
uint64_t cmib_owned_cell(uint64_t j, int32_t size, int32_t rank);
uint64_t cmib_owned_cell_count(uint64_t ncells, int32_t size, int32_t rank);
int cmib_comm_owned_cells(cmib_context *ctx, uint64_t *n_owned);
int cmib_download_cells_owned(cmib_context *ctx, double *number_density, double *temperature, double *ionic_fractions,
                              double *heating);
/* ... and the distributed upload: host arrays of the owned cells in the same order; cmib_comm_gather_owned_cells then
 * replicates the opacity records (all a shoot reads) — the metal fractions stay with the owner, who is the one that
 * updates them: every rank moves 1/size of the bytes and the gather carries 32 of the 128 bytes per cell */
int cmib_upload_cells_owned(cmib_context *ctx, const double *number_density, const double *temperature,
                            const double *ionic_fractions);
int cmib_comm_gather_owned_cells(cmib_context *ctx);

/* ---- measured ceilings of the part (roofline denominators) --------------- */
/* Scattered FP64 RED/s and scattered 16-byte gathers/s of THIS device on tables of `n_cells` records
 * (tools/microbench/red_bench.cu as a library call): every lane of a warp touches a different 128-byte line.
 * These are the L1TEX-lane ceilings the walk of an L2-resident grid runs against; bench.py measures them in
 * the run that reports them. */
int cmib_measure_scatter_rates(cmib_context *ctx, uint64_t n_cells, double *red_per_s, double *gather_per_s);

/* ---- test hooks (parity against the oracle on identical inputs) -------- */
/* CartesianDensityGrid::interact on explicit packets (src/CartesianDensityGrid.cpp:375-452).
 * pos/dir [np][3], sigma [np][14], sigma_He_corr/nu/weight/tau [np].  Accumulates
 * into the context's accumulators (full 16-term layout is forced).  Outputs:
 * final_pos [np][3], final_cell [np] (-1 = left the box), nsteps [np] (cells with
 * n > 0 that received a contribution), trace [np][max_trace] visit order of those
 * cells (-1 padded); trace may be NULL. */
int cmib_march_packets(cmib_context *ctx, int64_t np, const double *pos, const double *dir,
                       const double *sigma, const double *sigma_He_corr, const double *nu,
                       const double *weight, const double *tau, double *final_pos,
                       int64_t *final_cell, int32_t *nsteps, int32_t max_trace, int64_t *trace);
/* replaces: DensityGrid::integrate_optical_depth(const Photon&) (src/DensityGrid.hpp:392,
 * src/CartesianDensityGrid.cpp:328-363): optical depth from each packet's position to the edge of
 * the box along its direction, for np packets given as pos[np][3], dir[np][3], sigma_H[np] and
 * A_He*sigma_He[np]; same crossing order and arithmetic as the reference */
int cmib_integrate_optical_depth(cmib_context *ctx, int64_t np, const double *pos, const double *dir,
                                 const double *sigma_H, const double *sigma_He_corr, double *optical_depth);
/* PhotonSource::get_random_photon for packets [offset, offset+n): pos, dir [n][3],
 * nu [n], sigma [n][14], sigma_He_corr [n], tau [n] (the first optical depth). */
int cmib_sample_packets(cmib_context *ctx, int64_t n, uint64_t offset, uint64_t seed,
                        uint32_t iteration, double *pos, double *dir, double *nu, double *sigma,
                        double *sigma_He_corr, double *tau);
/* CrossSections::get_cross_section for the current model; sigma [n][14] */
int cmib_eval_cross_sections(cmib_context *ctx, int64_t n, const double *nu, double *sigma);
/* RecombinationRates::get_recombination_rate; alpha [n][14] */
int cmib_eval_recombination_rates(cmib_context *ctx, int64_t n, const double *T, double *alpha);
/* ChargeTransferRates; out [n][3][14] = recombination-with-H, ionization-with-H,
 * recombination-with-He as function of T4 = T/1e4 K */
int cmib_eval_charge_transfer(cmib_context *ctx, int64_t n, const double *T4, double *out);
/* LineCoolingData::get_cooling; abund [n][13] */
int cmib_eval_line_cooling(cmib_context *ctx, int64_t n, const double *T, const double *ne,
                           const double *abund, double *cooling);
/* LineCoolingData::solve_system_of_linear_equations; A [n][25], B [n][5] in/out */
int cmib_eval_solve5(cmib_context *ctx, int64_t n, double *A, double *B, int32_t *status);
/* set_reemission_probabilities; out [n][5] */
int cmib_eval_reemission_probabilities(cmib_context *ctx, int64_t n, const double *T, double *out);
/* IonizationStateCalculator::calculate_ionization_state(jfac, hfac, cell) on SoA cells:
 * J [14][n], heat [2][n], ndens [n], T [n] -> x [14][n], heat_out [2][n] */
int cmib_eval_ionization_state(cmib_context *ctx, int64_t n, double jfac, double hfac,
                               const double *J, const double *heat, const double *ndens,
                               const double *T, double *x, double *heat_out);
/* TemperatureCalculator::compute_cooling_and_heating_balance: j [n][14], h [n][2]
 * normalised; outputs h0, he0, gain, loss [n], metals [n][12] */
int cmib_eval_cooling_heating_balance(cmib_context *ctx, int64_t n, const double *T,
                                      const double *ndens, const double *j, const double *h,
                                      const double *midz, double *h0, double *he0, double *gain,
                                      double *loss, double *metals);
/* TemperatureCalculator::calculate_temperature(cell, jfac, hfac, midpoint): J [14][n],
 * heat [2][n], ndens, T, cr_factor (NULL = -1), midz (NULL = 0) -> T_out, x [14][n],
 * heat_out [2][n] */
int cmib_eval_temperature(cmib_context *ctx, int64_t n, double jfac, double hfac, const double *J,
                          const double *heat, const double *ndens, const double *T,
                          const double *cr_factor, const double *midz, double *T_out, double *x,
                          double *heat_out);
/* spectrum tables as built on the host for the device (compare with the reference's):
 * which 0 Planck [3][1000]; 1 H-Lyc, 2 He-Lyc: freq[1000] temp[100] cdf[100][1000];
 * 3 He-2-photon: freq[1000] cdf[1000].  Unused outputs may be NULL. */
int cmib_get_spectrum_tables(cmib_context *ctx, int which, double *a, double *b, double *c);
/* sample n frequencies on the device: which 0 source spectrum, 1 H-Lyc(T), 2 He-Lyc(T),
 * 3 He-2-photon, 4 spectrum of the continuous source */
int cmib_sample_spectrum(cmib_context *ctx, int which, double temperature, uint64_t seed,
                         int64_t n, double *nu);

#ifdef __cplusplus
}
#endif
#endif /* CMIB_H */
