/*
 * tables.cuh — read-only atomic-data tables.
 *
 * Each table exists twice: `d_<name>` in __constant__ memory for kernels (all
 * lanes of a warp read the same row at the same time, i.e. the constant-cache
 * broadcast case) and `h_<name>` as a host array for the host-side table
 * builders (re-emission spectra are tabulated on the host from the cross
 * sections, HydrogenLymanContinuumSpectrum.cpp:40-122).  CMIB_TBL(name) picks
 * the right one for the current compilation pass.
 */
#pragma once
#include "cmib_common.cuh"

namespace cmib {

#if defined(__CUDACC__)
#define CMIB_CONST_TABLE(type, name, dims) __constant__ type d_##name dims
#include "atomic_data.inc"
#include "linecooling_data.inc"
#undef CMIB_CONST_TABLE
#endif

#define CMIB_CONST_TABLE(type, name, dims) static const type h_##name dims
#include "atomic_data.inc"
#include "linecooling_data.inc"
#undef CMIB_CONST_TABLE

#if defined(__CUDA_ARCH__)
#define CMIB_TBL(name) d_##name
#else
#define CMIB_TBL(name) h_##name
#endif

} // namespace cmib
