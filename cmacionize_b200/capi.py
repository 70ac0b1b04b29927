"""ctypes binding of the C ABI in ``include/cmib.h`` (``libcmib.so``).

This is the only way Python reaches the product: there is no Python or CPU
implementation behind it.  If the shared library is missing the import fails
loudly; if no sm_100 GPU is present ``Context(...)`` fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

NUM_IONS = 14
NUM_HEAT = 2
NUM_ELEMENTS = 6
NUM_REEMIT = 5
NUM_PACKET_TYPES = 4

ION_NAMES = ["H_n", "He_n", "C_p1", "C_p2", "N_n", "N_p1", "N_p2", "O_n", "O_p1", "Ne_n", "Ne_p1",
             "S_p1", "S_p2", "S_p3"]

CROSS_SECTIONS_FIXED_VALUE, CROSS_SECTIONS_VERNER = 0, 1
RECOMBINATION_FIXED_VALUE, RECOMBINATION_VERNER = 0, 1
SPECTRUM_MONOCHROMATIC, SPECTRUM_PLANCK, SPECTRUM_UNIFORM, SPECTRUM_TABULATED = 0, 1, 2, 3
CONTINUOUS_NONE, CONTINUOUS_ISOTROPIC, CONTINUOUS_PLANAR, CONTINUOUS_DISTANT_STAR, CONTINUOUS_EXTENDED_DISC = 0, 1, 2, 3, 4
CONTINUOUS_SPIRAL_GALAXY = 5
REEMISSION_NONE, REEMISSION_PHYSICAL, REEMISSION_FIXED_VALUE = 0, 1, 2

# CMIB_LIB: load another build of the same ABI (A/B timing of kernel variants)
LIB_PATH = Path(os.environ.get("CMIB_LIB") or (Path(__file__).resolve().parent / "libcmib.so"))


class CmibError(RuntimeError):
    pass


class GridDesc(C.Structure):
    _fields_ = [("anchor", C.c_double * 3), ("sides", C.c_double * 3),
                ("ncell", C.c_int32 * 3), ("periodic", C.c_int32 * 3)]


class TemperatureParams(C.Structure):
    _fields_ = [("do_temperature_calculation", C.c_int32),
                ("minimum_number_of_iterations", C.c_uint32),
                ("epsilon_convergence", C.c_double),
                ("maximum_number_of_iterations", C.c_uint32),
                ("pah_heating_factor", C.c_double),
                ("cosmic_ray_heating_factor", C.c_double),
                ("cosmic_ray_heating_limit", C.c_double),
                ("cosmic_ray_heating_scale_length", C.c_double),
                ("minimum_ionized_temperature", C.c_double)]


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m cmacionize_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(str(LIB_PATH))
    lib.cmib_last_error.restype = C.c_char_p
    lib.cmib_kernel_launch_count.restype = C.c_uint64
    lib.cmib_distribute.restype = C.c_uint64
    lib.cmib_distribute.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    lib.cmib_owned_cell.restype = C.c_uint64
    lib.cmib_owned_cell.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    lib.cmib_owned_cell_count.restype = C.c_uint64
    lib.cmib_owned_cell_count.argtypes = [C.c_uint64, C.c_int32, C.c_int32]
    lib.cmib_distribute_block.restype = None
    lib.cmib_distribute_block.argtypes = [C.c_int32, C.c_int32, C.c_uint64, C.c_uint64, C.POINTER(C.c_uint64),
                                          C.POINTER(C.c_uint64)]
    if lib.cmib_abi_version() != 1:
        raise ImportError("libcmib.so ABI version mismatch")
    lib.cmib_set_abort_on_error(1 if os.environ.get("CMIB_ABORT_ON_ERROR") == "1" else 0)
    return lib


lib = _load()

_vp = C.c_void_p


def _p(a):
    """pointer to a C-contiguous numpy array (or NULL)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"], "array must be C contiguous"
    return a.ctypes.data_as(_vp)


def _f64(a, shape=None):
    if a is None:
        return None
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), f"expected shape {shape}, got {a.shape}"
    return a


def _check(rc: int) -> None:
    if rc != 0:
        raise CmibError(lib.cmib_last_error().decode())


def kernel_launch_count() -> int:
    return int(lib.cmib_kernel_launch_count())


def distribute(number: int, size: int, rank: int) -> int:
    """MPICommunicator::distribute (src/MPICommunicator.hpp:207-222); no GPU needed"""
    return int(lib.cmib_distribute(number, size, rank))


def distribute_block(rank: int, size: int, begin: int, end: int):
    """MPICommunicator::distribute_block (src/MPICommunicator.hpp:237-255); no GPU needed"""
    lo, hi = C.c_uint64(), C.c_uint64()
    lib.cmib_distribute_block(rank, size, begin, end, C.byref(lo), C.byref(hi))
    return int(lo.value), int(hi.value)


def owned_cell(j: int, size: int, rank: int) -> int:
    """cell index of work item j of rank `rank`: the multi-GPU state update deals chunks of 1024 cells round-robin"""
    return int(lib.cmib_owned_cell(j, size, rank))


def owned_cell_count(ncells: int, size: int, rank: int) -> int:
    return int(lib.cmib_owned_cell_count(ncells, size, rank))


def comm_unique_id() -> bytes:
    """128 bytes that rank 0 hands to the other ranks before Context.comm_init_rank"""
    buf = C.create_string_buffer(128)
    _check(lib.cmib_comm_unique_id(buf))
    return buf.raw


class Context:
    """One Cartesian density grid resident on one B200 (``cmib_context``)."""

    def __init__(self, anchor, sides, ncell, periodic=(False, False, False), device: int = 0):
        d = GridDesc()
        for k in range(3):
            d.anchor[k] = float(anchor[k])
            d.sides[k] = float(sides[k])
            d.ncell[k] = int(ncell[k])
            d.periodic[k] = 1 if periodic[k] else 0
        self.ncell = tuple(int(v) for v in ncell)
        self.ncells = self.ncell[0] * self.ncell[1] * self.ncell[2]
        self._h = _vp()
        _check(lib.cmib_create(C.byref(d), C.c_int(device), C.byref(self._h)))

    def close(self):
        if self._h:
            lib.cmib_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def synchronize(self):
        _check(lib.cmib_synchronize(self._h))

    # ---- grid state -----------------------------------------------------
    def upload_cells(self, number_density, temperature, ionic_fractions, cosmic_ray_factor=None):
        n = _f64(number_density).reshape(-1)
        T = _f64(temperature).reshape(-1)
        x = _f64(ionic_fractions).reshape(NUM_IONS, -1)
        assert n.size == self.ncells and T.size == self.ncells and x.shape[1] == self.ncells
        cr = _f64(cosmic_ray_factor)
        _check(lib.cmib_upload_cells(self._h, _p(n), _p(T), _p(x), _p(cr)))

    def download_cells(self):
        n = np.empty(self.ncells)
        T = np.empty(self.ncells)
        x = np.empty((NUM_IONS, self.ncells))
        heat = np.empty((NUM_HEAT, self.ncells))
        _check(lib.cmib_download_cells(self._h, _p(n), _p(T), _p(x), _p(heat)))
        return n, T, x, heat

    def download_cells_into(self, n, T, x, heat):
        _check(lib.cmib_download_cells(self._h, _p(n), _p(T), _p(x), _p(heat)))

    def download_accumulators(self):
        J = np.empty((NUM_IONS, self.ncells))
        heat = np.empty((NUM_HEAT, self.ncells))
        _check(lib.cmib_download_accumulators(self._h, _p(J), _p(heat)))
        return J, heat

    def reset_accumulators(self):
        _check(lib.cmib_reset_accumulators(self._h))

    # ---- plugins ----------------------------------------------------------
    def set_abundances(self, He=0., C_=0., N=0., O=0., Ne=0., S=0.):
        a = _f64([He, C_, N, O, Ne, S])
        self.abundances = a.copy()
        _check(lib.cmib_set_abundances(self._h, _p(a)))

    def set_cross_sections(self, kind, fixed=None):
        f = _f64(fixed, (NUM_IONS,)) if fixed is not None else None
        _check(lib.cmib_set_cross_sections(self._h, C.c_int(kind), _p(f)))

    def set_bimodal_cross_sections(self, frequency_limit, low, high):
        lo, hi = _f64(low, (NUM_IONS,)), _f64(high, (NUM_IONS,))
        _check(lib.cmib_set_bimodal_cross_sections(self._h, C.c_double(frequency_limit), _p(lo), _p(hi)))

    def set_recombination_rates(self, kind, fixed=None):
        f = _f64(fixed, (NUM_IONS,)) if fixed is not None else None
        _check(lib.cmib_set_recombination_rates(self._h, C.c_int(kind), _p(f)))

    def set_sources(self, positions, weights, total_luminosity):
        if positions is None or len(positions) == 0:  # PhotonSourceDistribution: None
            _check(lib.cmib_set_sources(self._h, C.c_int32(0), None, None, C.c_double(0.)))
            return
        pos = _f64(positions).reshape(-1, 3)
        w = _f64(weights).reshape(-1)
        assert pos.shape[0] == w.size
        _check(lib.cmib_set_sources(self._h, C.c_int32(w.size), _p(pos), _p(w),
                                    C.c_double(total_luminosity)))

    def set_spectrum_table(self, frequencies, cumulative_distribution, role=0):
        """A tabulated spectrum (FaucherGiguere, WMBasic, ... : include/cmib.h); role 0 = discrete sources,
        1 = continuous source."""
        f, c = _f64(frequencies).reshape(-1), _f64(cumulative_distribution).reshape(-1)
        assert f.size == c.size
        _check(lib.cmib_set_spectrum_table(self._h, C.c_int(role), C.c_int32(f.size), _p(f), _p(c)))

    def set_distant_star_position(self, position):
        pos = _f64(position).reshape(3)
        _check(lib.cmib_set_distant_star_position(self._h, _p(pos)))

    def set_extended_disc_geometry(self, normal_axis, origin, scale_height):
        _check(lib.cmib_set_extended_disc_geometry(self._h, C.c_int(normal_axis), C.c_double(origin), C.c_double(scale_height)))

    def set_spiral_galaxy_geometry(self, scale_length_stars, scale_height_stars, bulge_over_total_ratio):
        _check(lib.cmib_set_spiral_galaxy_geometry(self._h, C.c_double(scale_length_stars), C.c_double(scale_height_stars),
                                                   C.c_double(bulge_over_total_ratio)))

    def set_planar_source_geometry(self, normal_axis, intercept, anchor, sides):
        a, sd = _f64(anchor).reshape(2), _f64(sides).reshape(2)
        _check(lib.cmib_set_planar_source_geometry(self._h, C.c_int(normal_axis), C.c_double(intercept), _p(a), _p(sd)))

    def set_continuous_source(self, kind, luminosity=0., spectrum_kind=0, spectrum_param=0.):
        """IsotropicContinuousPhotonSource + its spectrum (include/cmib.h); luminosity = total surface
        area of the box x total flux of the spectrum."""
        _check(lib.cmib_set_continuous_source(self._h, C.c_int(kind), C.c_double(luminosity),
                                              C.c_int(spectrum_kind), C.c_double(spectrum_param)))

    def set_spectrum(self, kind, param):
        _check(lib.cmib_set_spectrum(self._h, C.c_int(kind), C.c_double(param)))

    def set_reemission(self, kind, probability=0.364, frequency=0.):
        _check(lib.cmib_set_reemission(self._h, C.c_int(kind), C.c_double(probability),
                                       C.c_double(frequency)))

    def set_temperature_params(self, do_temperature_calculation=False,
                               minimum_number_of_iterations=3, epsilon_convergence=1.e-3,
                               maximum_number_of_iterations=100, pah_heating_factor=0.,
                               cosmic_ray_heating_factor=0., cosmic_ray_heating_limit=0.75,
                               cosmic_ray_heating_scale_length=1.33333 * 3.086e19,
                               minimum_ionized_temperature=4000.):
        p = TemperatureParams(int(bool(do_temperature_calculation)), minimum_number_of_iterations,
                              epsilon_convergence, maximum_number_of_iterations,
                              pah_heating_factor, cosmic_ray_heating_factor,
                              cosmic_ray_heating_limit, cosmic_ray_heating_scale_length,
                              minimum_ionized_temperature)
        _check(lib.cmib_set_temperature_params(self._h, C.byref(p)))

    # ---- iteration --------------------------------------------------------
    def update_reemission_probabilities(self):
        _check(lib.cmib_update_reemission_probabilities(self._h))

    def shoot(self, n_packets, packet_offset=0, seed=42, iteration=0, want_counters=True):
        if want_counters:
            tw = C.c_double(0.)
            tc = (C.c_double * NUM_PACKET_TYPES)()
            _check(lib.cmib_shoot(self._h, C.c_uint64(n_packets), C.c_uint64(packet_offset),
                                  C.c_uint64(seed), C.c_uint32(iteration), C.byref(tw), tc))
            return tw.value, np.array(list(tc))
        _check(lib.cmib_shoot(self._h, C.c_uint64(n_packets), C.c_uint64(packet_offset),
                              C.c_uint64(seed), C.c_uint32(iteration), None, None))
        return None

    def set_shoot_algorithm(self, algorithm: int):
        """0 = wavefront pipeline (production), 1 = one-thread-per-packet kernel (A/B check)"""
        _check(lib.cmib_set_shoot_algorithm(self._h, C.c_int(algorithm)))

    def set_shoot_timing(self, on=True):
        _check(lib.cmib_set_shoot_timing(self._h, C.c_int(1 if on else 0)))

    def shoot_timing(self, want_adds=True):
        """(prepare ms, march ms, rounds, accumulator adds) of the last shoot / since the last reset"""
        a, b, d = C.c_double(), C.c_double(), C.c_double()
        r = C.c_uint64()
        _check(lib.cmib_shoot_timing(self._h, C.byref(a), C.byref(b), C.byref(r),
                                     C.byref(d) if want_adds else None))
        return a.value, b.value, int(r.value), d.value

    def set_packet_conventions(self, conventions):
        """0: IonizationSimulation (default), 1: TaskBasedIonizationSimulation (include/cmib.h)"""
        _check(lib.cmib_set_packet_conventions(self._h, C.c_int(conventions)))

    def shoot_overlap(self):
        """(lanes of the last shoot, ms during which emission and march kernels ran side by side)"""
        n, o = C.c_int32(), C.c_double()
        _check(lib.cmib_shoot_overlap(self._h, C.byref(n), C.byref(o)))
        return int(n.value), o.value

    def update_state(self, loop, totweight=0.):
        _check(lib.cmib_update_state(self._h, C.c_uint32(loop), C.c_double(totweight)))

    def shoot_statistics(self):
        a = C.c_double(0.)
        b = C.c_double(0.)
        _check(lib.cmib_shoot_statistics(self._h, C.byref(a), C.byref(b)))
        return a.value, b.value

    def shoot_optical_depth(self):
        t = C.c_double(0.)
        _check(lib.cmib_shoot_optical_depth(self._h, C.byref(t)))
        return t.value

    # ---- multi-GPU (include/cmib.h: the reference's MPICommunicator for this path) -------------
    def comm_init_rank(self, size: int, rank: int, unique_id: bytes):
        assert len(unique_id) == 128
        _check(lib.cmib_comm_init_rank(self._h, C.c_int32(size), C.c_int32(rank), C.c_char_p(unique_id)))

    def comm_finalize(self):
        _check(lib.cmib_comm_finalize(self._h))

    def comm_info(self):
        """(rank, size, first cell, one past the last cell of this rank's block)"""
        r, n = C.c_int32(), C.c_int32()
        lo, hi = C.c_uint64(), C.c_uint64()
        _check(lib.cmib_comm_info(self._h, C.byref(r), C.byref(n), C.byref(lo), C.byref(hi)))
        return int(r.value), int(n.value), int(lo.value), int(hi.value)

    def exchange_and_update(self, loop, allreduce=False):
        """sum the accumulators over the ranks, update this rank's cell block, gather the opacity records"""
        _check(lib.cmib_comm_exchange_and_update(self._h, C.c_uint32(loop), C.c_int(1 if allreduce else 0)))

    def exchange_timing(self):
        ms = (C.c_double * 3)()
        _check(lib.cmib_comm_exchange_timing(self._h, ms))
        return tuple(ms)

    def comm_gather_state(self):
        _check(lib.cmib_comm_gather_state(self._h))

    def comm_gather_cells_all(self):
        _check(lib.cmib_comm_gather_cells_all(self._h))

    def update_state_block(self, loop, totweight, cell_begin, cell_end):
        _check(lib.cmib_update_state_block(self._h, C.c_uint32(loop), C.c_double(totweight), C.c_uint64(cell_begin),
                                           C.c_uint64(cell_end)))

    def upload_cells_block(self, cell_begin, cell_end, n, T, x):
        """host arrays hold ONLY the block: n, T [nb], x [14][nb]"""
        _check(lib.cmib_upload_cells_block(self._h, C.c_uint64(cell_begin), C.c_uint64(cell_end), _p(n), _p(T), _p(x)))

    def download_cells_block_into(self, cell_begin, cell_end, n, T, x, heat):
        _check(lib.cmib_download_cells_block(self._h, C.c_uint64(cell_begin), C.c_uint64(cell_end), _p(n), _p(T), _p(x),
                                             _p(heat)))

    def owned_cells(self):
        """number of cells this rank updates (all of them without a communicator)"""
        n = C.c_uint64()
        _check(lib.cmib_comm_owned_cells(self._h, C.byref(n)))
        return int(n.value)

    def upload_cells_owned(self, n, T, x):
        """host arrays hold ONLY the owned cells in work-item order: n, T [n_owned], x [14][n_owned]"""
        _check(lib.cmib_upload_cells_owned(self._h, _p(n), _p(T), _p(x)))

    def comm_gather_owned_cells(self):
        _check(lib.cmib_comm_gather_owned_cells(self._h))

    def download_cells_owned_into(self, n, T, x, heat):
        """the cells this rank owns, in work-item order (capi.owned_cell): n, T [n_owned], x [14][n_owned], heat [2][n_owned]"""
        _check(lib.cmib_download_cells_owned(self._h, _p(n), _p(T), _p(x), _p(heat)))

    def measure_scatter_rates(self, n_cells=None):
        """(scattered FP64 RED/s, scattered 16-byte gathers/s) of this device, measured now"""
        a, b = C.c_double(), C.c_double()
        _check(lib.cmib_measure_scatter_rates(self._h, C.c_uint64(n_cells or self.ncells), C.byref(a), C.byref(b)))
        return a.value, b.value

    def accumulator_buffer(self):
        ptr = _vp()
        n = C.c_uint64()
        _check(lib.cmib_accumulator_buffer(self._h, C.byref(ptr), C.byref(n)))
        return ptr.value, int(n.value)

    def stream(self):
        s = _vp()
        _check(lib.cmib_stream(self._h, C.byref(s)))
        return s.value

    # ---- test hooks -------------------------------------------------------
    def march_packets(self, pos, direction, sigma, sigma_He_corr, nu, weight, tau, max_trace=0):
        pos = _f64(pos).reshape(-1, 3)
        np_ = pos.shape[0]
        direction = _f64(direction, (np_, 3))
        sigma = _f64(sigma, (np_, NUM_IONS))
        she = _f64(sigma_He_corr, (np_,))
        nu = _f64(nu, (np_,))
        w = _f64(weight, (np_,))
        tau = _f64(tau, (np_,))
        fpos = np.empty((np_, 3))
        fcell = np.empty(np_, dtype=np.int64)
        nsteps = np.empty(np_, dtype=np.int32)
        trace = np.empty((np_, max_trace), dtype=np.int64) if max_trace > 0 else None
        _check(lib.cmib_march_packets(self._h, C.c_int64(np_), _p(pos), _p(direction), _p(sigma),
                                      _p(she), _p(nu), _p(w), _p(tau), _p(fpos), _p(fcell),
                                      _p(nsteps), C.c_int32(max_trace), _p(trace)))
        return fpos, fcell, nsteps, trace

    def integrate_optical_depth(self, pos, direction, sigma_H, sigma_He_corr):
        """DensityGrid::integrate_optical_depth for explicit packets (include/cmib.h)"""
        pos, d = _f64(pos).reshape(-1, 3), _f64(direction).reshape(-1, 3)
        sh, she = _f64(sigma_H).reshape(-1), _f64(sigma_He_corr).reshape(-1)
        out = np.empty(sh.size)
        _check(lib.cmib_integrate_optical_depth(self._h, C.c_int64(sh.size), _p(pos), _p(d), _p(sh), _p(she), _p(out)))
        return out

    def sample_packets(self, n, offset=0, seed=42, iteration=0):
        pos = np.empty((n, 3)); d = np.empty((n, 3)); nu = np.empty(n)
        sigma = np.empty((n, NUM_IONS)); she = np.empty(n); tau = np.empty(n)
        _check(lib.cmib_sample_packets(self._h, C.c_int64(n), C.c_uint64(offset), C.c_uint64(seed),
                                       C.c_uint32(iteration), _p(pos), _p(d), _p(nu), _p(sigma),
                                       _p(she), _p(tau)))
        return dict(pos=pos, dir=d, nu=nu, sigma=sigma, sigma_He_corr=she, tau=tau)

    def eval_cross_sections(self, nu):
        nu = _f64(nu).reshape(-1)
        out = np.empty((nu.size, NUM_IONS))
        _check(lib.cmib_eval_cross_sections(self._h, C.c_int64(nu.size), _p(nu), _p(out)))
        return out

    def eval_recombination_rates(self, T):
        T = _f64(T).reshape(-1)
        out = np.empty((T.size, NUM_IONS))
        _check(lib.cmib_eval_recombination_rates(self._h, C.c_int64(T.size), _p(T), _p(out)))
        return out

    def eval_charge_transfer(self, T4):
        T4 = _f64(T4).reshape(-1)
        out = np.empty((T4.size, 3, NUM_IONS))
        _check(lib.cmib_eval_charge_transfer(self._h, C.c_int64(T4.size), _p(T4), _p(out)))
        return out

    def eval_line_cooling(self, T, ne, abund):
        T = _f64(T).reshape(-1)
        ne = _f64(ne, (T.size,))
        abund = _f64(abund, (T.size, 13))
        out = np.empty(T.size)
        _check(lib.cmib_eval_line_cooling(self._h, C.c_int64(T.size), _p(T), _p(ne), _p(abund), _p(out)))
        return out

    def eval_solve5(self, A, B):
        A = _f64(A).reshape(-1, 25).copy()
        B = _f64(B).reshape(-1, 5).copy()
        st = np.empty(A.shape[0], dtype=np.int32)
        _check(lib.cmib_eval_solve5(self._h, C.c_int64(A.shape[0]), _p(A), _p(B), _p(st)))
        return A, B, st

    def eval_reemission_probabilities(self, T):
        T = _f64(T).reshape(-1)
        out = np.empty((T.size, NUM_REEMIT))
        _check(lib.cmib_eval_reemission_probabilities(self._h, C.c_int64(T.size), _p(T), _p(out)))
        return out

    def eval_ionization_state(self, jfac, hfac, J, heat, ndens, T):
        J = _f64(J).reshape(NUM_IONS, -1)
        n = J.shape[1]
        heat = _f64(heat, (NUM_HEAT, n)); ndens = _f64(ndens, (n,)); T = _f64(T, (n,))
        x = np.empty((NUM_IONS, n)); ho = np.empty((NUM_HEAT, n))
        _check(lib.cmib_eval_ionization_state(self._h, C.c_int64(n), C.c_double(jfac),
                                              C.c_double(hfac), _p(J), _p(heat), _p(ndens), _p(T),
                                              _p(x), _p(ho)))
        return x, ho

    def eval_cooling_heating_balance(self, T, ndens, j, h, midz=None):
        T = _f64(T).reshape(-1)
        n = T.size
        ndens = _f64(ndens, (n,)); j = _f64(j, (n, NUM_IONS)); h = _f64(h, (n, NUM_HEAT))
        midz = _f64(midz, (n,)) if midz is not None else None
        h0 = np.empty(n); he0 = np.empty(n); gain = np.empty(n); loss = np.empty(n)
        metals = np.empty((n, 12))
        _check(lib.cmib_eval_cooling_heating_balance(self._h, C.c_int64(n), _p(T), _p(ndens), _p(j),
                                                     _p(h), _p(midz), _p(h0), _p(he0), _p(gain),
                                                     _p(loss), _p(metals)))
        return h0, he0, gain, loss, metals

    def eval_temperature(self, jfac, hfac, J, heat, ndens, T, cr_factor=None, midz=None):
        J = _f64(J).reshape(NUM_IONS, -1)
        n = J.shape[1]
        heat = _f64(heat, (NUM_HEAT, n)); ndens = _f64(ndens, (n,)); T = _f64(T, (n,))
        cr = _f64(cr_factor, (n,)) if cr_factor is not None else None
        mz = _f64(midz, (n,)) if midz is not None else None
        To = np.empty(n); x = np.empty((NUM_IONS, n)); ho = np.empty((NUM_HEAT, n))
        _check(lib.cmib_eval_temperature(self._h, C.c_int64(n), C.c_double(jfac), C.c_double(hfac),
                                         _p(J), _p(heat), _p(ndens), _p(T), _p(cr), _p(mz), _p(To),
                                         _p(x), _p(ho)))
        return To, x, ho

    def get_spectrum_tables(self, which):
        if which == 0:
            a = np.empty((3, 1000))
            _check(lib.cmib_get_spectrum_tables(self._h, 0, _p(a), None, None))
            return a
        if which in (1, 2):
            f = np.empty(1000); t = np.empty(100); c = np.empty((100, 1000))
            _check(lib.cmib_get_spectrum_tables(self._h, which, _p(f), _p(t), _p(c)))
            return f, t, c
        f = np.empty(1000); c = np.empty(1000)
        _check(lib.cmib_get_spectrum_tables(self._h, 3, _p(f), _p(c), None))
        return f, c

    def sample_spectrum(self, which, temperature, n, seed=1):
        nu = np.empty(n)
        _check(lib.cmib_sample_spectrum(self._h, C.c_int(which), C.c_double(temperature),
                                        C.c_uint64(seed), C.c_int64(n), _p(nu)))
        return nu
