"""ctypes binding of ``oracle/_ref/libcmi_ref.so`` — the UNMODIFIED reference
(CMacIonize) compiled by ``oracle/build_ref.py`` plus the probes of
``oracle/ref_harness.cpp``.

TEST INFRASTRUCTURE ONLY.  Allowed importers: ``tests/``,
``__graft_entry__.smoke()``, ``bench.py`` (cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

NUM_IONS = 14
LIB_PATH = Path(__file__).resolve().parent / "_ref" / "libcmi_ref.so"

_lib = None


def available() -> bool:
    return LIB_PATH.exists()


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise FileNotFoundError(f"{LIB_PATH} missing: run `python oracle/build_ref.py`")
        _lib = C.CDLL(str(LIB_PATH))
        _lib.cmi_ref_run_paramfile.restype = C.c_int64
        _lib.cmi_ref_convert.restype = C.c_double
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def _f(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None:
        assert a.shape == tuple(shape), (a.shape, shape)
    return a


def max_threads() -> int:
    return int(lib().cmi_ref_max_threads())


def verner_cross_sections(nu):
    nu = _f(nu).reshape(-1)
    out = np.empty((nu.size, NUM_IONS))
    lib().cmi_ref_verner_cross_sections(C.c_int64(nu.size), _p(nu), _p(out))
    return out


def verner_recombination_rates(T):
    T = _f(T).reshape(-1)
    out = np.empty((T.size, NUM_IONS))
    lib().cmi_ref_verner_recombination_rates(C.c_int64(T.size), _p(T), _p(out))
    return out


def charge_transfer(T4):
    T4 = _f(T4).reshape(-1)
    out = np.empty((T4.size, 3, NUM_IONS))
    lib().cmi_ref_charge_transfer(C.c_int64(T4.size), _p(T4), _p(out))
    return out


def linecooling_get_cooling(T, ne, abund):
    T = _f(T).reshape(-1)
    ne = _f(ne, (T.size,))
    abund = _f(abund, (T.size, 13))
    out = np.empty(T.size)
    lib().cmi_ref_linecooling_get_cooling(C.c_int64(T.size), _p(T), _p(ne), _p(abund), _p(out))
    return out


def solve5(A, B):
    A = _f(A).reshape(-1, 25).copy()
    B = _f(B).reshape(-1, 5).copy()
    st = np.empty(A.shape[0], dtype=np.int32)
    lib().cmi_ref_solve5(C.c_int64(A.shape[0]), _p(A), _p(B), _p(st))
    return A, B, st


def reemission_probabilities(T):
    T = _f(T).reshape(-1)
    out = np.empty((T.size, 5))
    lib().cmi_ref_reemission_probabilities(C.c_int64(T.size), _p(T), _p(out))
    return out


def interact(anchor, sides, ncell, periodic, cell_n, cell_xH, cell_xHe, pos, direction, sigma,
             sigma_He_corr, nu, weight, tau, max_trace=0, J=None, heat=None):
    anchor = _f(anchor, (3,)); sides = _f(sides, (3,))
    ncell = np.ascontiguousarray(ncell, dtype=np.int32)
    periodic = np.ascontiguousarray([1 if p else 0 for p in periodic], dtype=np.int32)
    nc = int(ncell[0]) * int(ncell[1]) * int(ncell[2])
    cell_n = _f(cell_n).reshape(-1); cell_xH = _f(cell_xH).reshape(-1)
    cell_xHe = _f(cell_xHe).reshape(-1)
    assert cell_n.size == nc
    pos = _f(pos).reshape(-1, 3)
    n = pos.shape[0]
    direction = _f(direction, (n, 3)); sigma = _f(sigma, (n, NUM_IONS))
    she = _f(sigma_He_corr, (n,)); nu = _f(nu, (n,)); w = _f(weight, (n,)); tau = _f(tau, (n,))
    accumulate = 0
    if J is None:
        J = np.zeros((NUM_IONS, nc)); heat = np.zeros((2, nc))
    else:
        accumulate = 1
    fpos = np.empty((n, 3)); fcell = np.empty(n, dtype=np.int64)
    nsteps = np.empty(n, dtype=np.int32)
    trace = np.empty((n, max_trace), dtype=np.int64) if max_trace > 0 else None
    lib().cmi_ref_interact(_p(anchor), _p(sides), _p(ncell), _p(periodic), _p(cell_n), _p(cell_xH),
                           _p(cell_xHe), C.c_int64(n), _p(pos), _p(direction), _p(sigma), _p(she),
                           _p(nu), _p(w), _p(tau), C.c_int(accumulate), _p(J), _p(heat), _p(fpos),
                           _p(fcell), _p(nsteps), C.c_int32(max_trace), _p(trace))
    return dict(J=J, heat=heat, final_pos=fpos, final_cell=fcell, nsteps=nsteps, trace=trace)


def ionization_state(jfac, hfac, abundances, rr_kind, rr_fixed, J, heat, ndens, T):
    J = _f(J).reshape(NUM_IONS, -1)
    n = J.shape[1]
    heat = _f(heat, (2, n)); ndens = _f(ndens, (n,)); T = _f(T, (n,))
    ab = _f(abundances, (6,))
    rf = _f(rr_fixed if rr_fixed is not None else np.zeros(NUM_IONS), (NUM_IONS,))
    x = np.empty((NUM_IONS, n)); ho = np.empty((2, n))
    lib().cmi_ref_ionization_state(C.c_int64(n), C.c_double(jfac), C.c_double(hfac), _p(ab),
                                   C.c_int(rr_kind), _p(rf), _p(J), _p(heat), _p(ndens), _p(T),
                                   _p(x), _p(ho))
    return x, ho


def h_he_state(alphaH, alphaHe, jH, jHe, nH, AHe, T):
    arrs = [_f(a).reshape(-1) for a in (alphaH, alphaHe, jH, jHe, nH, AHe, T)]
    n = arrs[0].size
    h0 = np.empty(n); he0 = np.empty(n)
    lib().cmi_ref_h_he_state(C.c_int64(n), *[_p(a) for a in arrs], _p(h0), _p(he0))
    return h0, he0


def cooling_heating_balance(T, ndens, j, h, abundances, pahfac=0., crfac=0., crscale=0., midz=None,
                            rr_kind=1, rr_fixed=None):
    T = _f(T).reshape(-1)
    n = T.size
    ndens = _f(ndens, (n,)); j = _f(j, (n, NUM_IONS)); h = _f(h, (n, 2))
    ab = _f(abundances, (6,))
    rf = _f(rr_fixed if rr_fixed is not None else np.zeros(NUM_IONS), (NUM_IONS,))
    mz = _f(midz, (n,)) if midz is not None else None
    h0 = np.empty(n); he0 = np.empty(n); gain = np.empty(n); loss = np.empty(n)
    metals = np.empty((n, 12))
    lib().cmi_ref_cooling_heating_balance(C.c_int64(n), _p(T), _p(ndens), _p(j), _p(h), _p(ab),
                                          C.c_double(pahfac), C.c_double(crfac),
                                          C.c_double(crscale), _p(mz), C.c_int(rr_kind), _p(rf),
                                          _p(h0), _p(he0), _p(gain), _p(loss), _p(metals))
    return h0, he0, gain, loss, metals


def temperature(jfac, hfac, abundances, J, heat, ndens, T, rr_kind=1, rr_fixed=None, pahfac=0.,
                crfac=0., crlim=0.75, crscale=1.33333 * 3.086e19, min_ionized_T=4000., epsilon=1.e-3,
                max_iterations=100, cr_factor=None, midz=None):
    J = _f(J).reshape(NUM_IONS, -1)
    n = J.shape[1]
    heat = _f(heat, (2, n)); ndens = _f(ndens, (n,)); T = _f(T, (n,))
    ab = _f(abundances, (6,))
    rf = _f(rr_fixed if rr_fixed is not None else np.zeros(NUM_IONS), (NUM_IONS,))
    tp = _f([pahfac, crfac, crlim, crscale, min_ionized_T, epsilon, float(max_iterations)])
    cr = _f(cr_factor, (n,)) if cr_factor is not None else None
    mz = _f(midz, (n,)) if midz is not None else None
    To = np.empty(n); x = np.empty((NUM_IONS, n)); ho = np.empty((2, n))
    lib().cmi_ref_temperature(C.c_int64(n), C.c_double(jfac), C.c_double(hfac), _p(ab),
                              C.c_int(rr_kind), _p(rf), _p(tp), _p(J), _p(heat), _p(ndens), _p(T),
                              _p(cr), _p(mz), _p(To), _p(x), _p(ho))
    return To, x, ho


def planck_tables(temperature):
    out = np.empty((3, 1000))
    lib().cmi_ref_planck_tables(C.c_double(temperature), _p(out))
    return out


def lyc_tables(which, xs_kind=1, xs_fixed=None):
    xf = _f(xs_fixed if xs_fixed is not None else np.zeros(NUM_IONS), (NUM_IONS,))
    f = np.empty(1000); t = np.empty(100); c = np.empty((100, 1000))
    lib().cmi_ref_lyc_tables(C.c_int(which), C.c_int(xs_kind), _p(xf), _p(f), _p(t), _p(c))
    return f, t, c


def he2pc_tables():
    f = np.empty(1000); c = np.empty(1000)
    lib().cmi_ref_he2pc_tables(_p(f), _p(c))
    return f, c


def sample_spectrum(which, temperature, n, seed=42):
    nu = np.empty(n)
    lib().cmi_ref_sample_spectrum(C.c_int(which), C.c_double(temperature), C.c_int(seed),
                                  C.c_int64(n), _p(nu))
    return nu


def isotropic_incoming(anchor, sides, n, seed=42):
    """(uniforms [n,5], positions [n,3], directions [n,3]) of IsotropicContinuousPhotonSource"""
    a = np.ascontiguousarray(anchor, dtype=np.float64)
    sd = np.ascontiguousarray(sides, dtype=np.float64)
    u, pos, d = np.empty((n, 5)), np.empty((n, 3)), np.empty((n, 3))
    lib().cmi_ref_isotropic_incoming(_p(a), _p(sd), C.c_int(seed), C.c_int64(n), _p(u), _p(pos), _p(d))
    return u, pos, d


def faucher_giguere(redshift, n, seed=42):
    """dict(freq, cdf, total_flux, uniforms, nu) of FaucherGiguerePhotonSourceSpectrum(redshift)"""
    u, nu = np.empty(n), np.empty(n)
    freq, cdf = np.empty(4096), np.empty(4096)
    flux = C.c_double(0.)
    m = lib().cmi_ref_tabulated_spectrum(C.c_int(0), C.c_double(redshift), C.c_int(seed), C.c_int64(n), _p(u), _p(nu),
                                         _p(freq), _p(cdf), C.c_int(4096), C.byref(flux))
    assert m > 0
    return dict(freq=freq[:m].copy(), cdf=cdf[:m].copy(), total_flux=flux.value, uniforms=u, nu=nu)


def uniform_spectrum(n, seed=42):
    u, nu = np.empty(n), np.empty(n)
    lib().cmi_ref_tabulated_spectrum(C.c_int(1), C.c_double(0.), C.c_int(seed), C.c_int64(n), _p(u), _p(nu),
                                     None, None, C.c_int(0), None)
    return u, nu


def random_stream(seed, n):
    out = np.empty(n)
    lib().cmi_ref_random_stream(C.c_int(seed), C.c_int64(n), _p(out))
    return out


def photon_source_distribution(paramfile, capacity=4096):
    """(positions [n,3], weights [n], total luminosity) of the reference's PhotonSourceDistributionFactory"""
    info, pos, w = np.zeros(2), np.empty((capacity, 3)), np.empty(capacity)
    n = lib().cmi_ref_photon_source_distribution(str(paramfile).encode(), _p(info), _p(pos), _p(w), C.c_int(capacity))
    assert 0 <= n <= capacity
    return pos[:n].copy(), w[:n].copy(), info[1]


def planar_incoming(axis, intercept, anchor, sides, n, seed=42):
    """(uniforms [n,4], positions [n,3], directions [n,3]) of PlanarContinuousPhotonSource"""
    a = np.ascontiguousarray(anchor, dtype=np.float64)
    sd = np.ascontiguousarray(sides, dtype=np.float64)
    u, pos, d = np.empty((n, 4)), np.empty((n, 3)), np.empty((n, 3))
    lib().cmi_ref_planar_incoming(C.c_int(axis), C.c_double(intercept), _p(a), _p(sd), C.c_int(seed), C.c_int64(n),
                                  _p(u), _p(pos), _p(d))
    return u, pos, d


def sph_mapping(paramfile, mapping_type, x, y, z, h, m, ncell, xH_cells, box=None):
    """(number density per cell, neutral fraction per particle) of the reference's SPHArrayInterface"""
    arrs = [np.ascontiguousarray(a, dtype=np.float64) for a in (x, y, z, h, m, xH_cells)]
    dens, nH = np.empty(ncell), np.empty(arrs[0].size)
    ba = bs = None
    if box is not None:
        ba, bs = (np.ascontiguousarray(b, dtype=np.float64) for b in box)
    L = lib()
    L.cmi_ref_sph_mapping.restype = C.c_int64
    n = L.cmi_ref_sph_mapping(str(paramfile).encode(), mapping_type.encode(), _p(ba) if ba is not None else None,
                              _p(bs) if bs is not None else None, C.c_int64(arrs[0].size), *[_p(a) for a in arrs[:5]],
                              C.c_int64(ncell), _p(dens), _p(arrs[5]), _p(nH))
    assert n == ncell, (n, ncell)
    return dens, nH


def integrate_optical_depth(anchor, sides, ncell, periodic, n, xH, xHe, pos, direction, sigma_H, sigma_He_corr):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (anchor, sides, n, xH, xHe, pos, direction, sigma_H, sigma_He_corr)]
    nc = np.ascontiguousarray(ncell, dtype=np.int32)
    per = np.ascontiguousarray(periodic, dtype=np.int32)
    npk = a[7].size
    out = np.empty(npk)
    lib().cmi_ref_integrate_optical_depth(_p(a[0]), _p(a[1]), _p(nc), _p(per), _p(a[2]), _p(a[3]), _p(a[4]), C.c_int64(npk),
                                          _p(a[5]), _p(a[6]), _p(a[7]), _p(a[8]), _p(out))
    return out


def distant_star_incoming(anchor, sides, star, n, seed=42):
    """(positions [n,3], directions [n,3], exposed area) of DistantStarContinuousPhotonSource with RandomGenerator(seed)"""
    a, sd, st = (np.ascontiguousarray(v, dtype=np.float64) for v in (anchor, sides, star))
    pos, d = np.empty((n, 3)), np.empty((n, 3))
    L = lib()
    L.cmi_ref_distant_star_incoming.restype = C.c_double
    area = L.cmi_ref_distant_star_incoming(_p(a), _p(sd), _p(st), C.c_int(seed), C.c_int64(n), _p(pos), _p(d))
    return pos, d, area


def extended_disc_incoming(anchor, sides, axis, origin, scale_height, n, seed=42):
    """(positions [n,3], directions [n,3]) of ExtendedDiscContinuousPhotonSource with RandomGenerator(seed); axis 'x'/'y'/'z'"""
    a, sd = (np.ascontiguousarray(v, dtype=np.float64) for v in (anchor, sides))
    pos, d = np.empty((n, 3)), np.empty((n, 3))
    lib().cmi_ref_extended_disc_incoming(_p(a), _p(sd), axis.encode(), C.c_double(origin), C.c_double(scale_height),
                                         C.c_int(seed), C.c_int64(n), _p(pos), _p(d))
    return pos, d


def spiral_galaxy_incoming(anchor, sides, r_stars, h_stars, B_over_T, n, seed=42):
    """(positions [n,3], directions [n,3]) of SpiralGalaxyContinuousPhotonSource with RandomGenerator(seed)"""
    a, sd = (np.ascontiguousarray(v, dtype=np.float64) for v in (anchor, sides))
    pos, d = np.empty((n, 3)), np.empty((n, 3))
    lib().cmi_ref_spiral_galaxy_incoming(_p(a), _p(sd), C.c_double(r_stars), C.c_double(h_stars), C.c_double(B_over_T),
                                         C.c_int(seed), C.c_int64(n), _p(pos), _p(d))
    return pos, d


def abundances(paramfile):
    out = np.empty(6)
    lib().cmi_ref_abundances(str(paramfile).encode(), _p(out))
    return out


def parameter_cross_sections(paramfile, nu):
    nu = np.ascontiguousarray(nu, dtype=np.float64).reshape(-1)
    out = np.empty((nu.size, 14))
    lib().cmi_ref_parameter_cross_sections(str(paramfile).encode(), C.c_int64(nu.size), _p(nu), _p(out))
    return out


def masked_spectrum(paramfile, role="PhotonSourceSpectrum"):
    freq, cdf = np.empty(65536), np.empty(65536)
    flux = C.c_double(0.)
    n = lib().cmi_ref_masked_spectrum(str(paramfile).encode(), role.encode(), _p(freq), _p(cdf), C.c_int(65536), C.byref(flux))
    assert n > 0
    return dict(freq=freq[:n].copy(), cdf=cdf[:n].copy(), total_flux=flux.value)


def random_photons(paramfile, n, seed=42):
    """n x PhotonSource::get_random_photon of the parameter file's source with RandomGenerator(seed)"""
    pos, d, nu = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
    sigma, she, w = np.empty((n, 14)), np.empty(n), np.empty(n)
    lib().cmi_ref_random_photons(str(paramfile).encode(), C.c_int(seed), C.c_int64(n), _p(pos), _p(d), _p(nu), _p(sigma),
                                 _p(she), _p(w))
    return dict(pos=pos, dir=d, nu=nu, sigma=sigma, sigma_He_corr=she, weight=w)


def reemit_sequence(paramfile, xH, xHe, T, nu_in, seed=42):
    """n x PhotonSource::reemit with RandomGenerator(seed) -> (new frequency or 0, packet type, new direction)"""
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (xH, xHe, T, nu_in)]
    n = a[0].size
    nu, typ, d = np.empty(n), np.empty(n, dtype=np.int32), np.empty((n, 3))
    lib().cmi_ref_reemit_sequence(str(paramfile).encode(), C.c_int(seed), C.c_int64(n), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]),
                                  _p(nu), _p(typ), _p(d))
    return nu, typ, d


def use_all_host_threads():
    """Let the reference's OpenMP regions use every core this process may run on, whatever OMP_NUM_THREADS
    says: `torch.distributed.run` exports OMP_NUM_THREADS=1 to its workers, and the reference clamps its
    thread count to omp_get_max_threads() (WorkEnvironment.hpp:63-76).  Acts on the libgomp the reference
    library is linked against.  Returns the thread count."""
    import os
    lib()
    n = len(os.sched_getaffinity(0))
    C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(n))
    return n


def gadget_kernel_sums(pos, m, h, rho, T, xH, periodic, box_sides, queries):
    """GadgetSnapshotDensityFunction::operator() composed from the reference's Octree + CubicSplineKernel on particle
    arrays (SI) -> (number density, temperature, neutral fraction or -1) at the query points"""
    f = lambda a: np.ascontiguousarray(a, dtype=np.float64)
    pos, m, h, rho, T, q = f(pos), f(m), f(h), f(rho), f(T), f(queries)
    xH = None if xH is None else f(xH)
    box = f(box_sides)
    out = np.empty((len(q), 3))
    lib().cmi_ref_gadget_kernel_sums(C.c_int64(m.size), _p(pos), _p(m), _p(h), _p(rho), _p(T), None if xH is None else _p(xH),
                                     C.c_int(1 if periodic else 0), _p(box), C.c_int64(len(q)), _p(q), _p(out))
    return out[:, 0].copy(), out[:, 1].copy(), out[:, 2].copy()


def convert(value, unit_from, unit_to):
    return float(lib().cmi_ref_convert(C.c_double(value), unit_from.encode(), unit_to.encode()))


def paramfile_query(path, queries):
    """queries: list of (kind, key, default) -> (list of value strings, used-values dump)"""
    q = "\n".join(f"{k}|{key}|{d}" for k, key, d in queries) + "\n"
    buf = C.create_string_buffer(1 << 17)
    lib().cmi_ref_paramfile_query(str(path).encode(), q.encode(), buf, C.c_int(1 << 17))
    text = buf.value.decode()
    head, dump = text.split("---\n", 1)
    return head.strip("\n").split("\n"), dump


def run_paramfile(path, ncells, num_threads=-1, verbose=False):
    """Run the reference IonizationSimulation; returns (fields[32][ncell], times): n, T, x[14], J[14], heat[2]."""
    fields = np.zeros((32, ncells))
    times = np.zeros(2)
    n = lib().cmi_ref_run_paramfile(str(path).encode(), C.c_int(num_threads),
                                    C.c_int(1 if verbose else 0), _p(fields), C.c_int64(ncells),
                                    _p(times))
    if n != ncells:
        raise RuntimeError(f"reference run returned {n} cells, expected {ncells}")
    return fields, times


def run_paramfile_taskbased(path, ncells, num_threads=-1, verbose=False):
    """Run the reference TaskBasedIonizationSimulation (`CMacIonize --task-based`); fields[32][ncell] in the
    Cartesian cell order of the whole box: n, T, x[14], J[14] / abundance, heat[2] as the last temperature step left them."""
    fields = np.zeros((32, ncells))
    L = lib()
    L.cmi_ref_run_paramfile_taskbased.restype = C.c_int64
    n = L.cmi_ref_run_paramfile_taskbased(str(path).encode(), C.c_int(num_threads), C.c_int(1 if verbose else 0),
                                          _p(fields), C.c_int64(ncells))
    if n != ncells:
        raise RuntimeError(f"reference task-based run returned {n} cells, expected {ncells}")
    return fields


class Simulation:
    """The reference IonizationSimulation driven one iteration at a time (probe
    cmi_ref_sim_*: the loop body of IonizationSimulation::run on the reference's objects)."""

    def __init__(self, paramfile, num_threads=-1):
        L = lib()
        L.cmi_ref_sim_create.restype = C.c_void_p
        L.cmi_ref_sim_number_of_cells.restype = C.c_int64
        self._h = C.c_void_p(L.cmi_ref_sim_create(str(paramfile).encode(), C.c_int(num_threads)))
        self.ncells = int(L.cmi_ref_sim_number_of_cells(self._h))
        self.threads = int(L.cmi_ref_sim_threads(self._h))

    def iteration(self, loop, numphoton):
        out = np.zeros(8)
        lib().cmi_ref_sim_iteration(self._h, C.c_uint32(loop), C.c_uint64(int(numphoton)), _p(out))
        return dict(shoot_s=out[0], update_s=out[1], totweight=out[2], typecount=out[3:7].copy(),
                    prep_s=out[7])

    def fields(self):
        f = np.empty((32, self.ncells))
        lib().cmi_ref_sim_get_fields(self._h, _p(f))
        return f

    def set_state(self, n, T, x):
        f = np.empty((16, self.ncells))
        f[0] = n
        f[1] = T
        f[2:16] = x
        lib().cmi_ref_sim_set_state(self._h, _p(f))

    def close(self):
        if self._h:
            lib().cmi_ref_sim_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
