"""ctypes binding of the C++ host layer (``libcmih.so``: host/IonizationSimulation.hpp behind
host/host_api.cpp) — the reference-facing surface: a parameter file in, an
``IonizationSimulation`` that is initialised and run, fields out.

Nothing here computes: the host layer configures a ``cmib_context`` and orders the C-ABI calls of
one iteration; all physics runs in the CUDA library."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import capi

LIB_PATH = Path(__file__).resolve().parent / "libcmih.so"
if not LIB_PATH.exists():
    raise ImportError(f"{LIB_PATH} is missing: build it with `python -m cmacionize_b200.build`")
lib = C.CDLL(str(LIB_PATH))
lib.cmih_last_error.restype = C.c_char_p

QUANTITY = {"length": 6, "number_density": 8, "reaction_rate": 9, "surface_area": 10, "temperature": 11,
            "frequency": 5}


class HostError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise HostError(lib.cmih_last_error().decode())


def convert(value, unit_from, unit_to):
    out = C.c_double()
    _check(lib.cmih_convert(C.c_double(value), unit_from.encode(), unit_to.encode(), C.byref(out)))
    return out.value


def random_stream(seed, n):
    """n deviates of the host layer's RandomGenerator (the reference's RANLUX stream)"""
    out = np.empty(n)
    _check(lib.cmih_random_stream(C.c_int32(seed), C.c_int64(n), out.ctypes.data_as(C.c_void_p)))
    return out


def test_rendezvous(nthreads, rounds, failing_thread=-1, failing_round=-1):
    """per iteration: 1 = every device thread saw success, 0 = every thread saw the failure, -1 = they disagree"""
    out = np.zeros(rounds, dtype=np.int32)
    _check(lib.cmih_test_rendezvous(C.c_int(nthreads), C.c_int(rounds), C.c_int(failing_thread), C.c_int(failing_round),
                                    out.ctypes.data_as(C.c_void_p)))
    return out


def sph_mapping(paramfile, mapping_type, x, y, z, h, m, ncell, xH_cells, box=None):
    """(number density per cell, neutral fraction per particle) of the host layer's SPHArrayInterface: the
    device-free half of the cmi_* C ABI (host/SPHArrayInterface.hpp)"""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, h, m, xH_cells)]
    dens, nH = np.empty(ncell), np.empty(arrs[0].size)
    vp = C.c_void_p
    ba = bs = None
    if box is not None:
        ba, bs = (np.ascontiguousarray(b, dtype=np.float64) for b in box)
    _check(lib.cmih_sph_mapping(str(paramfile).encode(), mapping_type.encode(),
                                ba.ctypes.data_as(vp) if ba is not None else None,
                                bs.ctypes.data_as(vp) if bs is not None else None, C.c_int64(arrs[0].size),
                                *[a.ctypes.data_as(vp) for a in arrs[:5]], C.c_int64(ncell), dens.ctypes.data_as(vp),
                                arrs[5].ctypes.data_as(vp), nH.ctypes.data_as(vp)))
    return dens, nH


class ParameterFile:
    def __init__(self, filename):
        self._h = C.c_void_p()
        _check(lib.cmih_paramfile_open(str(filename).encode(), C.byref(self._h)))

    def close(self):
        if self._h:
            lib.cmih_paramfile_close(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        self.close()

    def get_string(self, key, default=""):
        buf = C.create_string_buffer(4096)
        _check(lib.cmih_paramfile_get_string(self._h, key.encode(), default.encode(), buf, 4096))
        return buf.value.decode()

    def get_double(self, key, default):
        out = C.c_double()
        _check(lib.cmih_paramfile_get_double(self._h, key.encode(), C.c_double(default), C.byref(out)))
        return out.value

    def get_physical(self, quantity, key, default):
        out = C.c_double()
        _check(lib.cmih_paramfile_get_physical(self._h, QUANTITY[quantity], key.encode(), default.encode(),
                                               C.byref(out)))
        return out.value

    def used_values(self):
        buf = C.create_string_buffer(1 << 16)
        _check(lib.cmih_paramfile_used_values(self._h, buf, 1 << 16))
        return buf.value.decode()

    def photon_source_spectrum(self, role="PhotonSourceSpectrum"):
        """dict(kind, param, total_flux, freq, cdf) of the file's spectrum for `role`"""
        info, freq, cdf = np.zeros(4), np.empty(65536), np.empty(65536)
        vp = C.c_void_p
        _check(lib.cmih_photon_source_spectrum(self._h, role.encode(), info.ctypes.data_as(vp), freq.ctypes.data_as(vp),
                                               cdf.ctypes.data_as(vp), C.c_int(65536)))
        m = int(info[3])
        return dict(kind=int(info[0]), param=info[1], total_flux=info[2], freq=freq[:m].copy(), cdf=cdf[:m].copy())

    def sample_spectrum(self, n, seed=42, role="PhotonSourceSpectrum"):
        """n frequencies sampled on the host with RandomGenerator(seed) from the file's spectrum"""
        nu = np.empty(n)
        _check(lib.cmih_sample_spectrum(self._h, role.encode(), C.c_int32(seed), C.c_int64(n), nu.ctypes.data_as(C.c_void_p)))
        return nu

    def photon_source_distribution(self, capacity=4096):
        """(positions [n,3], weights [n], total luminosity) of the file's PhotonSourceDistribution"""
        info, pos, w = np.zeros(2), np.empty((capacity, 3)), np.empty(capacity)
        vp = C.c_void_p
        _check(lib.cmih_photon_source_distribution(self._h, info.ctypes.data_as(vp), pos.ctypes.data_as(vp),
                                                   w.ctypes.data_as(vp), C.c_int(capacity)))
        n = int(info[0])
        return pos[:n].copy(), w[:n].copy(), info[1]

    def initial_number_density(self, ncells):
        """number density per cell after DensityFunction + DensityMask on the file's grid (no device needed)"""
        dens = np.empty(ncells)
        _check(lib.cmih_initial_number_density(self._h, C.c_int64(ncells), dens.ctypes.data_as(C.c_void_p)))
        return dens

    def write_snapshot(self, output_folder, iteration, number_density, temperature, ionic_fractions, time=0.):
        """the file's DensityGridWriter (Gadget HDF5 by default, AsciiFile) on given cell arrays (no device
        needed) -> name of the snapshot file"""
        n = np.ascontiguousarray(number_density, dtype=np.float64).reshape(-1)
        T = np.ascontiguousarray(temperature, dtype=np.float64).reshape(-1)
        x = np.ascontiguousarray(ionic_fractions, dtype=np.float64).reshape(14, -1)
        assert T.size == n.size and x.shape[1] == n.size
        name = C.create_string_buffer(4096)
        _check(lib.cmih_write_snapshot(self._h, str(output_folder).encode(), C.c_uint32(iteration), C.c_double(time),
                                       C.c_int64(n.size), n.ctypes.data_as(C.c_void_p), T.ctypes.data_as(C.c_void_p),
                                       x.ctypes.data_as(C.c_void_p), name, C.c_int(4096)))
        return name.value.decode()

    def initial_grid(self, ncells):
        """(number density, temperature, neutral H fraction) per cell after DensityFunction + DensityMask on the file's
        grid: the path IonizationSimulation::initialize takes (whole-grid fills included), no device needed"""
        dens, temp, xH = np.empty(ncells), np.empty(ncells), np.empty(ncells)
        vp = C.c_void_p
        _check(lib.cmih_initial_grid(self._h, C.c_int64(ncells), dens.ctypes.data_as(vp), temp.ctypes.data_as(vp),
                                     xH.ctypes.data_as(vp)))
        return dens, temp, xH

    def abundances(self):
        out = np.empty(6)
        _check(lib.cmih_abundances(self._h, out.ctypes.data_as(C.c_void_p)))
        return out

    def cross_sections(self, nu):
        """sigma [n,14] of the file's FixedValue / Bimodal CrossSections at the frequencies nu"""
        nu = np.ascontiguousarray(nu, dtype=np.float64).reshape(-1)
        out = np.empty((nu.size, 14))
        _check(lib.cmih_parameter_cross_sections(self._h, C.c_int64(nu.size), nu.ctypes.data_as(C.c_void_p),
                                                 out.ctypes.data_as(C.c_void_p)))
        return out

    def density_function(self, x):
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1, 3)
        n = x.shape[0]
        dens, temp, xH = np.empty(n), np.empty(n), np.empty(n)
        vp = C.c_void_p
        _check(lib.cmih_density_function(self._h, C.c_int64(n), x.ctypes.data_as(vp), dens.ctypes.data_as(vp),
                                         temp.ctypes.data_as(vp), xH.ctypes.data_as(vp)))
        return dens, temp, xH


class IonizationSimulation:
    """cmi::IonizationSimulation (host/IonizationSimulation.hpp) on one GPU."""

    def __init__(self, parameterfile, device=0, write_output=False, verbose=False, ngpus=1, task_based=False):
        self._h = C.c_void_p()
        self.ngpus = ngpus
        create = lib.cmih_simulation_create_task_based if task_based else lib.cmih_simulation_create_multi
        _check(create(str(parameterfile).encode(), C.c_int(device), C.c_int(ngpus), C.c_int(1 if write_output else 0),
                      C.c_int(1 if verbose else 0), C.byref(self._h)))
        info = (C.c_double * 4)()
        _check(lib.cmih_simulation_info(self._h, info))
        self.ncells = int(info[0])
        self.number_of_iterations = int(info[1])
        self.number_of_photons = int(info[2])
        self.total_luminosity = float(info[3])

    def close(self):
        if self._h:
            lib.cmih_simulation_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def initialize(self):
        _check(lib.cmih_simulation_initialize(self._h))

    def run(self):
        _check(lib.cmih_simulation_run(self._h))

    def iteration(self, loop, numphoton):
        out = (C.c_double * 8)()
        _check(lib.cmih_simulation_iteration(self._h, C.c_uint32(loop), C.c_uint64(int(numphoton)), out))
        return dict(totweight=out[0], typecount=np.array(out[1:5]), shoot_s=out[5], update_s=out[6])

    def fields(self, device_index=0):
        """(number density, temperature, ionic fractions [14][ncell], heating [2][ncell])"""
        ctx = C.c_void_p()
        _check(lib.cmih_simulation_gather_state(self._h))   # no-op on one device
        _check(lib.cmih_simulation_context_of(self._h, C.c_int(device_index), C.byref(ctx)))
        n = np.empty(self.ncells); T = np.empty(self.ncells)
        x = np.empty((capi.NUM_IONS, self.ncells)); heat = np.empty((capi.NUM_HEAT, self.ncells))
        vp = C.c_void_p
        rc = capi.lib.cmib_download_cells(ctx, n.ctypes.data_as(vp), T.ctypes.data_as(vp), x.ctypes.data_as(vp),
                                          heat.ctypes.data_as(vp))
        if rc != 0:
            raise HostError(capi.lib.cmib_last_error().decode())
        return n, T, x, heat


def write_particle_snapshot(filename, pos, m, h, rho, T=None, xH=None, periodic=-1, boxsize=(0., 0., 0.),
                            units_cgs=(0., 0., 0.), unit_time_cgs=1., time=0., sfr=None, stars=None):
    """a Gadget-style SPH snapshot from particle arrays (host/HDF5Writer.hpp): input for the GadgetSnapshot
    density function and source distribution.  periodic < 0: no /RuntimePars group; units_cgs = (0, 0, 0): no /Units
    group; stars = (positions, formation times, masses) -> /PartType4; sfr -> /PartType0/StarFormationRate"""
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    pos, m, h, rho, T, xH, sfr = f(pos), f(m), f(h), f(rho), f(T), f(xH), f(sfr)
    sp, sf, sm = (f(a) for a in stars) if stars is not None else (None, None, None)
    vp = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    box = np.ascontiguousarray(boxsize, dtype=np.float64)
    _check(lib.cmih_write_particle_snapshot(str(filename).encode(), C.c_int64(m.size), vp(pos), vp(m), vp(h), vp(rho), vp(T),
                                            vp(xH), C.c_int(periodic), vp(box), C.c_double(units_cgs[0]),
                                            C.c_double(units_cgs[1]), C.c_double(units_cgs[2]), C.c_double(unit_time_cgs),
                                            C.c_double(time), vp(sfr), C.c_int64(0 if sm is None else sm.size), vp(sp), vp(sf),
                                            vp(sm)))


class HDF5Input:
    """host/HDF5Reader.hpp through the C probes of libcmih (tests)."""

    def __init__(self, filename):
        self.filename = str(filename).encode()

    def exists(self, path):
        out = C.c_int(0)
        _check(lib.cmih_hdf5_exists(self.filename, path.encode(), C.byref(out)))
        return bool(out.value)

    def attribute_names(self, path):
        buf = C.create_string_buffer(1 << 18)
        _check(lib.cmih_hdf5_attribute_names(self.filename, path.encode(), buf, C.c_int(1 << 18)))
        return [l for l in buf.value.decode().split("\n") if l]

    def string_attribute(self, path, name):
        buf = C.create_string_buffer(1 << 16)
        n = C.c_int(0)
        _check(lib.cmih_hdf5_attribute(self.filename, path.encode(), name.encode(), 0, buf, C.c_int(1 << 16), None, 0, C.byref(n)))
        return buf.value.decode()

    def numeric_attribute(self, path, name):
        v = np.empty(64)
        n = C.c_int(0)
        _check(lib.cmih_hdf5_attribute(self.filename, path.encode(), name.encode(), 1, None, 0,
                                       v.ctypes.data_as(C.c_void_p), 64, C.byref(n)))
        return v[:n.value].copy()

    def dataset(self, path):
        dims = (C.c_int64 * 4)()
        ndim, count = C.c_int(0), C.c_int64(0)
        _check(lib.cmih_hdf5_dataset(self.filename, path.encode(), None, C.c_int64(0), dims, C.byref(ndim), C.byref(count)))
        out = np.empty(count.value)
        _check(lib.cmih_hdf5_dataset(self.filename, path.encode(), out.ctypes.data_as(C.c_void_p), C.c_int64(count.value), dims,
                                     C.byref(ndim), C.byref(count)))
        return out.reshape([dims[k] for k in range(ndim.value)])
