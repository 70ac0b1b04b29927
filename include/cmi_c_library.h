/*
 * cmi_c_library.h — the reference's coarse C ABI, served by the B200 backend (libcmih.so).
 *
 * Declarations are those of /root/reference/c/cmi_c_library.h:31-56 (implementation
 * src/CMILibrary.cpp:48-208): an SPH code hands over particle arrays and gets neutral fractions
 * back.  A program linked against the reference's libCMILibrary can link against libcmih.so
 * instead, unchanged.  Differences, all on the host side of the call:
 *   - num_thread is ignored; the GPU is chosen with the environment variable CMIB_DEVICE (default 0);
 *   - mapping_type: "M_over_V" and "centroid" (cmacionize_b200/host/SPHArrayInterface.hpp); "Petkova"
 *     aborts with an explanatory message;
 *   - errors print a message and abort(), like cmac_error.
 * Units, array ownership (the caller's) and the global-singleton life cycle
 * (cmi_init* ... cmi_compute_neutral_fraction_* ... cmi_destroy) are the reference's.
 */
#ifndef CMI_C_LIBRARY_H
#define CMI_C_LIBRARY_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* c/cmi_c_library.h:31-33 — non-periodic box (the SimulationBox of the parameter file) */
void cmi_init(const char *parameter_file, const int num_thread, const double unit_length_in_SI,
              const double unit_mass_in_SI, const char *mapping_type, const int talk);
/* c/cmi_c_library.h:34-38 — periodic box given in the caller's length unit, double precision */
void cmi_init_periodic_dp(const char *parameter_file, const int num_thread, const double unit_length_in_SI,
                          const double unit_mass_in_SI, const double *box_anchor, const double *box_sides,
                          const char *mapping_type, const int talk);
/* c/cmi_c_library.h:39-43 — the same with single precision box */
void cmi_init_periodic_sp(const char *parameter_file, const int num_thread, const double unit_length_in_SI,
                          const double unit_mass_in_SI, const float *box_anchor, const float *box_sides,
                          const char *mapping_type, const int talk);
/* c/cmi_c_library.h:44 */
void cmi_destroy();

/* c/cmi_c_library.h:46-55 — positions, smoothing lengths, masses (caller's units) -> neutral fractions;
 * double, mixed and single precision */
void cmi_compute_neutral_fraction_dp(const double *x, const double *y, const double *z, const double *h,
                                     const double *m, double *nH, const size_t N);
void cmi_compute_neutral_fraction_mp(const double *x, const double *y, const double *z, const float *h,
                                     const float *m, float *nH, const size_t N);
void cmi_compute_neutral_fraction_sp(const float *x, const float *y, const float *z, const float *h,
                                     const float *m, float *nH, const size_t N);

#ifdef __cplusplus
}
#endif

#endif /* CMI_C_LIBRARY_H */
