"""CPU tier: the product's physics headers (cmacionize_b200/csrc/*.cuh), compiled
for the host by tests/hostcheck, must agree BIT FOR BIT with the compiled
reference on the same inputs — same libm, no FMA contraction on either side, so
any difference is a logic error.  The GPU tier repeats these checks on the
device through the C ABI (test_gpu_*.py)."""
import ctypes as C

import numpy as np
import pytest

from cases import check_wall_intersection_scenarios, wall_intersection_scenarios, ABUNDANCES, MARCH_GRIDS, march_case, state_cells


def p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def test_verner_cross_sections_bitexact(hostcheck, ref):
    rng = np.random.default_rng(1)
    nu = 3.288e15 * np.exp(rng.uniform(-0.1, np.log(6.), 20000))
    out = np.empty((nu.size, 14))
    hostcheck.hc_verner_cross_sections(C.c_int64(nu.size), p(nu), p(out))
    assert np.array_equal(out, ref.verner_cross_sections(nu))


def test_recombination_charge_transfer_reemission_bitexact(hostcheck, ref):
    rng = np.random.default_rng(2)
    T = np.exp(rng.uniform(np.log(50.), np.log(1e7), 20000))
    out = np.empty((T.size, 14))
    hostcheck.hc_verner_recombination_rates(C.c_int64(T.size), p(T), p(out))
    assert np.array_equal(out, ref.verner_recombination_rates(T))
    T4 = T * 1e-4
    out = np.empty((T.size, 3, 14))
    hostcheck.hc_charge_transfer(C.c_int64(T.size), p(T4), p(out))
    assert np.array_equal(out, ref.charge_transfer(T4))
    out = np.empty((T.size, 5))
    hostcheck.hc_reemission_probabilities(C.c_int64(T.size), p(T), p(out))
    assert np.array_equal(out, ref.reemission_probabilities(T))


def test_line_cooling_and_solver_bitexact(hostcheck, ref):
    rng = np.random.default_rng(3)
    n = 5000
    T = np.exp(rng.uniform(np.log(3000.), np.log(40000.), n))
    ne = np.exp(rng.uniform(np.log(1e3), np.log(1e12), n))
    ne[:10] = 0.  # get_cooling returns 1e-99 (LineCoolingData.cpp:1771-1775)
    ab = rng.uniform(0, 1e-4, (n, 13))
    out = np.empty(n)
    hostcheck.hc_line_cooling(C.c_int64(n), p(T), p(ne), p(ab), p(out))
    assert np.array_equal(out, ref.linecooling_get_cooling(T, ne, ab))
    A = rng.uniform(-1, 1, (10000, 25))
    B = rng.uniform(-1, 1, (10000, 5))
    A[:5] = 0.  # singular systems must be reported, not solved
    A2, B2 = A.copy(), B.copy()
    st = np.empty(10000, dtype=np.int32)
    hostcheck.hc_solve5(C.c_int64(10000), p(A2), p(B2), p(st))
    Ar, Br, sr = ref.solve5(A, B)
    assert np.array_equal(st, sr) and (st[:5] == 1).all()
    assert np.array_equal(B2[5:], Br[5:])


def test_spectrum_tables_bitexact(hostcheck, ref):
    for temp in (20000., 40000.):
        o = np.empty((3, 1000))
        hostcheck.hc_planck_tables(C.c_double(temp), p(o))
        assert np.array_equal(o, ref.planck_tables(temp))
    fixed = np.zeros(14)
    fixed[0], fixed[1] = 6.3e-22, 1e-22
    for which in (0, 1):
        for kind, xf in ((1, np.zeros(14)), (0, fixed)):
            f = np.empty(1000); t = np.empty(100); c = np.empty((100, 1000))
            hostcheck.hc_lyc_tables(C.c_int(which), C.c_int(kind), p(xf), p(f), p(t), p(c))
            rf, rt, rc = ref.lyc_tables(which, kind, xf)
            assert np.array_equal(f, rf) and np.array_equal(t, rt) and np.array_equal(c, rc)
    f = np.empty(1000); c = np.empty(1000)
    hostcheck.hc_he2pc_tables(p(f), p(c))
    rf, rc = ref.he2pc_tables()
    assert np.array_equal(f, rf) and np.array_equal(c, rc)


def test_ionization_state_bitexact(hostcheck, ref, golden):
    J, heat, nd, T = state_cells(golden, reps=20)
    n = nd.size
    for ab, kind, fixed in ((ABUNDANCES, 1, np.zeros(14)),
                            (np.zeros(6), 0, np.array([4e-19] + [0.] * 13))):
        x = np.empty((14, n)); ho = np.empty((2, n))
        hostcheck.hc_ionization_state(C.c_int64(n), C.c_double(1.3), C.c_double(1.3 * 6.626e-34), p(ab),
                                      C.c_int(kind), p(fixed), p(J), p(heat), p(nd), p(T), p(x), p(ho))
        xr, hr = ref.ionization_state(1.3, 1.3 * 6.626e-34, ab, kind, fixed, J, heat, nd, T)
        assert np.array_equal(x, xr, equal_nan=True)
        assert np.array_equal(ho, hr)


def test_cooling_heating_balance_bitexact(hostcheck, ref, golden):
    g = golden["ioneng"]
    n = len(g)
    T = np.ascontiguousarray(g[:, 16]); nd = np.ascontiguousarray(g[:, 19]) * 1e6
    j = np.ascontiguousarray(g[:, :14]); h = np.ascontiguousarray(g[:, 14:16]) * 1e-7
    mz = np.linspace(-1e19, 1e19, n)
    for pah, cr, scale in ((1., 0., 0.75), (0.3, 0.4, 2e19)):
        h0 = np.empty(n); he0 = np.empty(n); gain = np.empty(n); loss = np.empty(n)
        met = np.empty((n, 12))
        hostcheck.hc_cooling_heating_balance(C.c_int64(n), p(T), p(nd), p(j), p(h), p(ABUNDANCES),
                                             C.c_double(pah), C.c_double(cr), C.c_double(scale), p(mz),
                                             C.c_int(1), p(np.zeros(14)), p(h0), p(he0), p(gain),
                                             p(loss), p(met))
        r = ref.cooling_heating_balance(T, nd, j, h, ABUNDANCES, pah, cr, scale, midz=mz)
        for a, b in zip((h0, he0, gain, loss, met), r):
            assert np.array_equal(a, b)


def test_temperature_bitexact(hostcheck, ref, golden):
    J, heat, nd, T = state_cells(golden, reps=10)
    n = nd.size
    rng = np.random.default_rng(11)
    crf = rng.uniform(-1, 2, n)
    mz = rng.uniform(-1e19, 1e19, n)
    for tp in ([0., 0., 0.75, 1.33333 * 3.086e19, 4000., 1e-3, 100.],
               [0.5, 0.2, 0.75, 1e19, 4000., 1e-3, 100.]):
        tpa = np.array(tp)
        To = np.empty(n); x = np.empty((14, n)); ho = np.empty((2, n))
        hostcheck.hc_temperature(C.c_int64(n), C.c_double(1.), C.c_double(1.), p(ABUNDANCES), C.c_int(1),
                                 p(np.zeros(14)), p(tpa), p(J), p(heat), p(nd), p(T), p(crf), p(mz),
                                 p(To), p(x), p(ho))
        Tr, xr, hr = ref.temperature(1., 1., ABUNDANCES, J, heat, nd, T, pahfac=tp[0], crfac=tp[1],
                                     crlim=tp[2], crscale=tp[3], cr_factor=crf, midz=mz)
        assert np.array_equal(To, Tr)
        assert np.array_equal(x, xr)
        assert np.array_equal(ho, hr)
        assert 500. in Tr and Tr.max() <= 30000.


@pytest.mark.parametrize("name", list(MARCH_GRIDS))
def test_march_bitexact(hostcheck, ref, name):
    """Voxel traversal: visited-cell sequence, final position, final cell, and (serial
    accumulation order being identical) even the J/heating sums are bit-identical."""
    npk = 4000
    c = march_case(name, npk)
    nc = int(np.prod(c["ncell"]))
    mt = 512
    r = ref.interact(c["anchor"], c["sides"], c["ncell"], c["periodic"], c["n"], c["xH"], c["xHe"],
                     c["pos"], c["dir"], c["sigma"], c["sigma_He_corr"], c["nu"], c["weight"],
                     c["tau"], max_trace=mt)
    J = np.zeros((14, nc)); heat = np.zeros((2, nc)); fp = np.empty((npk, 3))
    fc = np.empty(npk, np.int64); ns = np.empty(npk, np.int32); tr = np.empty((npk, mt), np.int64)
    hostcheck.hc_march_packets(p(c["anchor"]), p(c["sides"]), p(c["ncell"]), p(c["periodic"]), p(c["n"]),
                               p(c["xH"]), p(c["xHe"]), C.c_int64(npk), p(c["pos"]), p(c["dir"]),
                               p(c["sigma"]), p(c["sigma_He_corr"]), p(c["nu"]), p(c["weight"]),
                               p(c["tau"]), p(J), p(heat), p(fp), p(fc), p(ns), C.c_int32(mt), p(tr))
    assert np.array_equal(ns, r["nsteps"])
    assert np.array_equal(tr, r["trace"])
    assert np.array_equal(fc, r["final_cell"])
    assert np.array_equal(fp, r["final_pos"])
    assert np.array_equal(J, r["J"]) and np.array_equal(heat, r["heat"])
    assert (ns > 0).any()


def test_wall_intersection_scenarios_of_the_reference_unit_test(hostcheck, ref):
    """test/testCartesianDensityGrid.cpp:310-465: the product's walk (host build) and the compiled
    reference on the nine get_wall_intersection scenarios."""
    from cases import check_wall_intersection_scenarios, wall_intersection_scenarios
    c = wall_intersection_scenarios()
    npk, nc, mt = len(c["tau"]), 16 ** 3, 4
    r = ref.interact(c["anchor"], c["sides"], c["ncell"], c["periodic"], c["n"], c["xH"], c["xHe"], c["pos"], c["dir"],
                     c["sigma"], c["sigma_He_corr"], c["nu"], c["weight"], c["tau"], max_trace=mt)
    check_wall_intersection_scenarios(c, r["final_pos"], r["final_cell"], r["nsteps"], r["trace"], r["J"])
    J = np.zeros((14, nc)); heat = np.zeros((2, nc)); fp = np.empty((npk, 3))
    fc = np.empty(npk, np.int64); ns = np.empty(npk, np.int32); tr = np.empty((npk, mt), np.int64)
    hostcheck.hc_march_packets(p(c["anchor"]), p(c["sides"]), p(c["ncell"]), p(c["periodic"]), p(c["n"]),
                               p(c["xH"]), p(c["xHe"]), C.c_int64(npk), p(c["pos"]), p(c["dir"]),
                               p(c["sigma"]), p(c["sigma_He_corr"]), p(c["nu"]), p(c["weight"]),
                               p(c["tau"]), p(J), p(heat), p(fp), p(fc), p(ns), C.c_int32(mt), p(tr))
    check_wall_intersection_scenarios(c, fp, fc, ns, tr, J)
    assert np.array_equal(fp, r["final_pos"]) and np.array_equal(tr, r["trace"])


def test_isotropic_continuous_source_bitexact(hostcheck, ref):
    """IsotropicContinuousPhotonSource::get_random_incoming_direction fed with the reference
    generator's own deviates: start position on the box surface and direction, bit for bit
    (boxes off the origin, non-cubic; the start must lie in the half-open box)."""
    for anchor, sides in (([-5., -5., -5.], [10., 10., 10.]), ([1e17, -3e17, 2e16], [2e17, 5e17, 3e17])):
        u, pos, d = ref.isotropic_incoming(anchor, sides, 20000, seed=7)
        pos2, d2 = np.empty_like(pos), np.empty_like(d)
        a, sd = np.array(anchor), np.array(sides)
        hostcheck.hc_isotropic_incoming(p(a), p(sd), C.c_int64(len(u)), p(np.ascontiguousarray(u)), p(pos2), p(d2))
        assert np.array_equal(d2, d)
        assert np.array_equal(pos2, pos)
        assert (pos2 >= a).all() and (pos2 < a + sd).all()
        on_face = (np.isclose(pos2, a, rtol=0, atol=1e-12 * sd) | np.isclose(pos2, a + sd, rtol=0, atol=1e-12 * sd)).any(axis=1)
        assert on_face.all()
        inward = np.where(np.isclose(pos2, a, rtol=0, atol=1e-12 * sd), d2, np.where(np.isclose(pos2, a + sd, rtol=0, atol=1e-12 * sd), -d2, 1.))
        assert (inward >= 0.).all()


def test_tabulated_and_uniform_spectra_bitexact(hostcheck, ref):
    """The sampling rule shared by the reference's tabulated spectra (FaucherGiguere here, z = 0 and
    z = 7: FaucherGiguerePhotonSourceSpectrum.cpp:234-247) and the Uniform spectrum, fed with the
    reference generator's own deviates and the reference object's own two arrays: bit for bit."""
    for z in (0., 7.):
        d = ref.faucher_giguere(z, 50000, seed=5)
        nu = np.empty(50000)
        hostcheck.hc_tabulated_frequency(C.c_int32(d["freq"].size), p(d["freq"]), p(d["cdf"]), C.c_int64(nu.size),
                                         p(d["uniforms"]), p(nu))
        assert np.array_equal(nu, d["nu"])
        assert d["cdf"][0] == 0. and d["cdf"][-1] == 1. and (np.diff(d["cdf"]) >= 0).all()
    u, nu_ref = ref.uniform_spectrum(50000, seed=6)
    nu = np.empty(50000)
    hostcheck.hc_tabulated_frequency(C.c_int32(0), None, None, C.c_int64(nu.size), p(u), p(nu))
    assert np.array_equal(nu, nu_ref)


def test_planar_continuous_source_bitexact(hostcheck, ref):
    """PlanarContinuousPhotonSource::get_random_incoming_direction for the three normal axes, fed with
    the reference generator's own deviates: position on the rectangle and direction, bit for bit."""
    for axis in range(3):
        anchor, sides, intercept = np.array([-1e17, 2e16]), np.array([3e17, 5e16]), 1.5e16 * (axis - 1)
        u, pos, d = ref.planar_incoming(axis, intercept, anchor, sides, 20000, seed=11 + axis)
        pos2, d2 = np.empty_like(pos), np.empty_like(d)
        hostcheck.hc_planar_incoming(C.c_int(axis), C.c_double(intercept), p(anchor), p(sides), C.c_int64(len(u)),
                                     p(np.ascontiguousarray(u)), p(pos2), p(d2))
        assert np.array_equal(pos2, pos) and np.array_equal(d2, d)
        assert (pos2[:, axis] == intercept).all()


@pytest.mark.parametrize("grid", ["unit16", "stromgren64_corner", "vacuum_holes", "single_cell", "noncubic_periodic_xz"])
def test_integrate_optical_depth_bitexact(hostcheck, ref, grid):
    """CartesianDensityGrid::integrate_optical_depth (optical depth to the edge of the box) on explicit
    packets: the host build of march.cuh against the reference, bit for bit (corner starts, axis-aligned
    and diagonal rays, vacuum cells; on the periodic grid only rays that leave through the y faces —
    the others never end in the reference either)."""
    c = march_case(grid, 3000)
    pos, d = c["pos"], c["dir"]
    if c["periodic"].any():
        keep = np.abs(d[:, 1]) > 0.2
        pos, d = np.ascontiguousarray(pos[keep]), np.ascontiguousarray(d[keep])
    sh = np.ascontiguousarray(c["sigma"][: len(pos), 0])
    she = np.ascontiguousarray(c["sigma_He_corr"][: len(pos)])
    r = ref.integrate_optical_depth(c["anchor"], c["sides"], c["ncell"], c["periodic"], c["n"], c["xH"], c["xHe"],
                                    pos, d, sh, she)
    out = np.empty(len(pos))
    hostcheck.hc_integrate_optical_depth(p(c["anchor"]), p(c["sides"]), p(c["ncell"]), p(c["periodic"]), p(c["n"]),
                                         p(c["xH"]), p(c["xHe"]), C.c_int64(len(pos)), p(pos), p(d), p(sh), p(she), p(out))
    assert np.array_equal(out, r)
    assert (r >= 0).all() and (r > 0).mean() > 0.9


def test_distant_star_continuous_source_bitexact(hostcheck, ref):
    """DistantStarContinuousPhotonSource::get_random_incoming_direction (rejection sampling, a variable
    number of deviates per packet) driven by the RANLUX stream on both sides: every start position and
    direction bit for bit — one, two and three exposed faces, below and above the box."""
    anchor, sides = np.array([-1e17, -2e17, -1.5e17]), np.array([2e17, 4e17, 3e17])
    for star in ([-6e17, 0., 0.], [3e17, 5e17, 0.5e17], [-4e17, -7e17, 6e17], [0.2e17, 0.3e17, 9e17]):
        st = np.array(star)
        pos, d, area = ref.distant_star_incoming(anchor, sides, st, 5000, seed=9)
        pos2, d2 = np.empty_like(pos), np.empty_like(d)
        hostcheck.hc_distant_star_incoming(p(anchor), p(sides), p(st), C.c_int(9), C.c_int64(len(pos)), p(pos2), p(d2))
        assert np.array_equal(d2, d) and np.array_equal(pos2, pos)
        exposed = (st < anchor) | (st > anchor + sides)
        assert area == sum(sides[(k + 1) % 3] * sides[(k + 2) % 3] for k in range(3) if exposed[k])
        on_face = (np.isclose(pos2, anchor, rtol=0, atol=1e-9 * sides) | np.isclose(pos2, anchor + sides, rtol=0, atol=1e-9 * sides))
        assert (on_face & exposed).any(axis=1).all()


def test_extended_disc_continuous_source_bitexact(hostcheck, ref):
    """ExtendedDiscContinuousPhotonSource::get_random_incoming_direction (a Gaussian height redrawn while it falls
    outside the box: a variable number of deviates per packet) driven by the RANLUX stream on both sides: every
    start position and direction bit for bit — every axis, discs inside, off centre and beyond the box."""
    anchor, sides = np.array([-1e17, -2e17, -1.5e17]), np.array([2e17, 4e17, 3e17])
    for axis, origin, height in ((2, 0., 0.4e17), (0, 0.6e17, 1e17), (1, -2.5e17, 0.5e17), (2, 0.2e17, 8e17)):
        pos, d = ref.extended_disc_incoming(anchor, sides, "xyz"[axis], origin, height, 5000, seed=11)
        pos2, d2 = np.empty_like(pos), np.empty_like(d)
        hostcheck.hc_extended_disc_incoming(p(anchor), p(sides), C.c_int(axis), C.c_double(origin), C.c_double(height),
                                            C.c_int(11), C.c_int64(len(pos)), p(pos2), p(d2))
        assert np.array_equal(d2, d) and np.array_equal(pos2, pos)
        assert ((pos2 >= anchor) & (pos2 <= anchor + sides)).all()
        assert np.unique(pos2[:, axis]).size > 4000 and abs(np.linalg.norm(d2, axis=1) - 1.).max() < 1e-15


def test_spiral_galaxy_continuous_source_bitexact(hostcheck, ref):
    """SpiralGalaxyContinuousPhotonSource::get_random_incoming_direction (bulge or disc positions redrawn until one
    lies in the box, the disc radius from a tabulated cumulative luminosity) driven by the RANLUX stream on both sides:
    every start position and direction bit for bit — a box around the whole galaxy, a thin slab, an off-centre box
    that rejects most of the disc, bulge-only and disc-only mixtures."""
    KPC = 3.086e19
    for anchor, sides, rs, hs, bt in (([-12., -12., -12.], [24., 24., 24.], 5., 0.6, 0.2),
                                      ([-6., -6., -0.4], [12., 12., 0.8], 3., 0.3, 0.5),
                                      ([-1., -2., -3.], [9., 4., 5.], 5., 0.6, 0.2),
                                      ([-4., -4., -4.], [8., 8., 8.], 2., 0.2, 0.),
                                      ([-4., -4., -4.], [8., 8., 8.], 2., 0.2, 1.)):
        a, sd = np.array(anchor) * KPC, np.array(sides) * KPC
        pos, d = ref.spiral_galaxy_incoming(a, sd, rs * KPC, hs * KPC, bt, 5000, seed=13)
        pos2, d2, tables = np.empty_like(pos), np.empty_like(d), np.empty(2002)
        hostcheck.hc_spiral_galaxy_incoming(p(a), p(sd), C.c_double(rs * KPC), C.c_double(hs * KPC), C.c_double(bt), C.c_int(13),
                                            C.c_int64(len(pos)), p(pos2), p(d2), p(tables))
        assert np.array_equal(d2, d) and np.array_equal(pos2, pos)
        assert ((pos2 >= a) & (pos2 < a + sd)).all() and abs(np.linalg.norm(d2, axis=1) - 1.).max() < 1e-15
        assert tables[1000] == 1.2 * np.sqrt((a ** 2).sum()) and tables[2001] == 1. and (np.diff(tables[1001:]) >= 0).all()


def _source_paramfile(tmp_path, continuous):
    yml = tmp_path / "sources.yml"
    yml.write_text("number of sources: 3\nsource[0]:\n  position: [0. pc, 0. pc, 0. pc]\n  luminosity: 2.e49 s^-1\n"
                   "source[1]:\n  position: [1. pc, -2. pc, 0.5 pc]\n  luminosity: 1.e49 s^-1\n"
                   "source[2]:\n  position: [-3. pc, 3. pc, -1. pc]\n  luminosity: 1.e49 s^-1\n")
    text = ("SimulationBox:\n  anchor: [-5. pc, -4. pc, -3. pc]\n  sides: [10. pc, 8. pc, 6. pc]\n  periodicity: [false, false, false]\n"
            f"PhotonSourceDistribution:\n  type: AsciiFile\n  filename: {yml}\n"
            "PhotonSourceSpectrum:\n  type: Planck\n  temperature: 40000. K\n"
            "AbundanceModel:\n  type: FixedValue\n  He: 0.1\n"
            "CrossSections:\n  type: Verner\nDiffuseReemissionHandler:\n  type: Physical\n")
    if continuous:
        text += ("ContinuousPhotonSource:\n  type: Isotropic\nContinuousPhotonSourceSpectrum:\n  type: Planck\n"
                 "  temperature: 25000. K\n  ionizing flux: 1.e14 m^-2 s^-1\n")
    pf = tmp_path / f"source_{int(continuous)}.param"
    pf.write_text(text)
    return pf


@pytest.mark.parametrize("continuous", [False, True])
def test_emission_with_the_reference_stream_is_bitexact(hostcheck, ref, tmp_path, continuous):
    """PhotonSource::get_random_photon (PhotonSource.cpp:208-249) with the reference's own random stream:
    the device's emission code (emit_primary + packet_cross_sections, source.cuh), compiled for the host and
    fed by the RANLUX generator, returns the reference's packets — source pick, direction, Planck frequency,
    14 Verner cross sections, weight — bit for bit; with an isotropic continuous source mixed in (half of the
    packets, other spectrum, other weight, five more deviates) as well.  The production path replaces the
    stream by per-packet Philox streams; this pins everything else."""
    pf = _source_paramfile(tmp_path, continuous)
    n = 20000
    r = ref.random_photons(pf, n, seed=31)
    PC = 3.086e16
    anchor, sides = np.array([-5 * PC, -4 * PC, -3 * PC]), np.array([10 * PC, 8 * PC, 6 * PC])
    src = np.array([[0., 0., 0.], [1 * PC, -2 * PC, 0.5 * PC], [-3 * PC, 3 * PC, -1 * PC]])
    ref_src, ref_w, ref_L = ref.photon_source_distribution(pf)
    assert np.array_equal(ref_src, src)
    area = 2 * (sides[0] * sides[1] + sides[0] * sides[2] + sides[1] * sides[2])
    Lc = area * 1e14 if continuous else 0.
    pos, d, nu = np.empty((n, 3)), np.empty((n, 3)), np.empty(n)
    sig, she, w = np.empty((n, 14)), np.empty(n), np.empty(n)
    hostcheck.hc_random_photons(p(anchor), p(sides), C.c_int(3), p(np.ascontiguousarray(ref_src)), p(ref_w), C.c_double(ref_L),
                                C.c_int(1), C.c_double(40000.), C.c_double(Lc), C.c_double(25000.), C.c_double(0.1),
                                C.c_int(31), C.c_int64(n), p(pos), p(d), p(nu), p(sig), p(she), p(w))
    assert np.array_equal(pos, r["pos"]) and np.array_equal(d, r["dir"])
    assert np.array_equal(nu, r["nu"])
    assert np.array_equal(sig, r["sigma"]) and np.array_equal(she, r["sigma_He_corr"])
    assert np.array_equal(w, r["weight"])
    if continuous:
        assert 0.45 < (w != 1.).mean() < 0.55 and np.unique(w).size == 2
    else:
        assert (w == 1.).all() and abs(np.all(pos == 0., axis=1).mean() - 0.5) < 0.02


def test_reemission_with_the_reference_stream_is_bitexact(hostcheck, ref, tmp_path):
    """PhotonSource::reemit + PhysicalDiffuseReemissionHandler::reemit (PhotonSource.cpp:272-308,
    PhysicalDiffuseReemissionHandler.cpp:219-370) with the reference's own random stream: absorbed by H or He,
    channel, new frequency from the H-Lyc / He-Lyc / two-photon tables, new direction — 50000 absorptions
    over a wide range of cells and frequencies, every outcome bit for bit (a different number of deviates per
    branch: one wrong branch would derail the rest of the sequence)."""
    pf = _source_paramfile(tmp_path, False)
    rng = np.random.default_rng(8)
    n = 50000
    xH = np.exp(rng.uniform(np.log(1e-5), 0., n))
    xHe = np.exp(rng.uniform(np.log(1e-5), 0., n))
    T = np.exp(rng.uniform(np.log(3000.), np.log(25000.), n))
    nu_in = 3.288465385e15 * np.exp(rng.uniform(np.log(1.001), np.log(3.9), n))
    r_nu, r_type, r_dir = ref.reemit_sequence(pf, xH, xHe, T, nu_in, seed=77)
    nu, typ, d = np.empty(n), np.empty(n, dtype=np.int32), np.empty((n, 3))
    hostcheck.hc_reemit_sequence(C.c_double(0.1), C.c_int(77), C.c_int64(n), p(xH), p(xHe), p(T), p(nu_in), p(nu), p(typ), p(d))
    assert np.array_equal(typ, r_type)
    assert np.array_equal(nu, r_nu) and np.array_equal(d, r_dir)
    assert set(np.unique(typ)) == {1, 2, 3} and 0.2 < (nu > 0).mean() < 0.8      # H-diffuse, He-diffuse, absorbed


def test_whole_shoot_with_the_reference_stream_is_bitexact(hostcheck, ref, tmp_path):
    """One complete shoot of the reference (IonizationSimulation iteration 0, one thread:
    IonizationPhotonShootJob::execute, IonizationPhotonShootJob.hpp:117-146) against the product's
    shoot_packet logic (shoot.cuh) compiled for the host and fed by the same RANLUX stream: emission,
    optical depth, voxel walk, accumulation, Physical re-emission chain of 20000 packets on a 12^3 grid
    with Planck + Verner + He — the 14 mean-intensity accumulators of every cell, the total weight and the
    packet-type counters come out bit for bit.  With this, the only parts of the production path that are
    not bit-identical to the reference are the random stream (per-packet Philox, by design), the order of the
    atomic adds and CUDA's libm."""
    PC = 3.086e16
    nc = 12
    pf = tmp_path / "shoot.param"
    pf.write_text(
        "SimulationBox:\n  anchor: [-3. pc, -3. pc, -3. pc]\n  sides: [6. pc, 6. pc, 6. pc]\n  periodicity: [false, false, false]\n"
        f"DensityGrid:\n  type: Cartesian\n  number of cells: [{nc}, {nc}, {nc}]\n"
        "DensityFunction:\n  type: Homogeneous\n  density: 100. cm^-3\n  temperature: 8000. K\n  neutral fraction H: 2.e-3\n"
        "PhotonSourceDistribution:\n  type: SingleStar\n  position: [0.2 pc, -0.1 pc, 0.3 pc]\n  luminosity: 1.e49 s^-1\n"
        "PhotonSourceSpectrum:\n  type: Planck\n  temperature: 40000. K\n"
        "AbundanceModel:\n  type: FixedValue\n  He: 0.1\n  C: 2.2e-4\n  N: 4.e-5\n  O: 3.3e-4\n  Ne: 5.e-5\n  S: 9.e-6\n"
        "CrossSections:\n  type: Verner\nRecombinationRates:\n  type: Verner\nDiffuseReemissionHandler:\n  type: Physical\n"
        "TemperatureCalculator:\n  do temperature calculation: false\n"
        f"IonizationSimulation:\n  number of photons: 20000\n  number of iterations: 1\n  random seed: 123\n  output folder: {tmp_path}\n")
    sim = ref.Simulation(pf, num_threads=1)
    f0 = sim.fields()
    # a state with real opacity in H and He, the same on both sides
    rng = np.random.default_rng(4)
    x = f0[2:16].copy()
    x[0] = np.exp(rng.uniform(np.log(1e-5), np.log(2e-3), nc ** 3))
    x[1] = np.exp(rng.uniform(np.log(1e-5), np.log(1e-2), nc ** 3))
    T = rng.uniform(6000., 12000., nc ** 3)
    sim.set_state(f0[0], T, x)
    npk = 20000
    out = sim.iteration(0, npk)
    f1 = sim.fields()
    sim.close()
    cells = np.ascontiguousarray(np.stack([f0[0], x[0], x[1], T], 1))
    anchor, sides = np.array([-3 * PC] * 3), np.array([6 * PC] * 3)
    ncell, per = np.array([nc] * 3, np.int32), np.zeros(3, np.int32)
    ip = np.array([1, 1, 1, 1, 0, 1], np.int32)      # 1 source, Planck, Verner, Physical, full layout, RANLUX stream
    dp = np.array([40000., 0.1, 0., 0.])
    sp, sw = np.array([[0.2 * PC, -0.1 * PC, 0.3 * PC]]), np.array([1.])
    acc = np.zeros(16 + 16 * nc ** 3)
    hostcheck.hc_shoot(p(anchor), p(sides), p(ncell), p(per), p(cells), p(ip), p(dp), p(sp), p(sw), None, C.c_uint64(npk),
                       C.c_uint64(0), C.c_uint64(123), C.c_uint32(0), p(acc))
    assert acc[0] == out["totweight"] == npk
    assert np.array_equal(acc[1:5], out["typecount"])
    assert acc[6] > 1.2 * npk                         # re-emissions happened
    slot = [0, 4, 8, 15, 3, 9, 13, 2, 11, 6, 12, 7, 10, 14]   # acc_slot() of the 14 ions (shoot.cuh)
    rec = acc[16:].reshape(nc ** 3, 16)
    for ion in range(14):
        assert np.array_equal(rec[:, slot[ion]], f1[16 + ion]), ion
    assert (f1[16] > 0).mean() > 0.9 and out["typecount"][0] + out["typecount"][1] + out["typecount"][2] > 0.05 * npk   # some escape too


def test_whole_simulation_with_the_reference_stream_is_bitexact(hostcheck, ref, tmp_path):
    """The reference's IonizationSimulation, one thread, 6 iterations of 20000 packets on a 10^3 Lexington-type
    grid (Planck star, Verner cross sections and rates, He and metals, Physical diffuse field, the temperature
    solve with line cooling from iteration 4 on) against the same loop composed from the product's physics
    headers on the host, fed by the reference's own random stream: the final temperature and all 14 ionic
    fractions of every cell are the reference's, bit for bit.  (State update arithmetic uses pow() in the host
    build, as the reference does; the device uses exp(a ln x), see test_gpu_physics.py for its bounds.)"""
    PC = 3.086e16
    nc = 10
    pf = tmp_path / "sim.param"
    pf.write_text(
        "SimulationBox:\n  anchor: [-3. pc, -3. pc, -3. pc]\n  sides: [6. pc, 6. pc, 6. pc]\n  periodicity: [false, false, false]\n"
        f"DensityGrid:\n  type: Cartesian\n  number of cells: [{nc}, {nc}, {nc}]\n"
        "DensityFunction:\n  type: Homogeneous\n  density: 100. cm^-3\n  temperature: 8000. K\n"
        "PhotonSourceDistribution:\n  type: SingleStar\n  position: [0.1 pc, 0.2 pc, -0.1 pc]\n  luminosity: 1.e49 s^-1\n"
        "PhotonSourceSpectrum:\n  type: Planck\n  temperature: 40000. K\n"
        "AbundanceModel:\n  type: FixedValue\n  He: 0.1\n  C: 2.2e-4\n  N: 4.e-5\n  O: 3.3e-4\n  Ne: 5.e-5\n  S: 9.e-6\n"
        "CrossSections:\n  type: Verner\nRecombinationRates:\n  type: Verner\nDiffuseReemissionHandler:\n  type: Physical\n"
        "TemperatureCalculator:\n  do temperature calculation: true\n"
        f"IonizationSimulation:\n  number of photons: 20000\n  number of iterations: 6\n  random seed: 321\n  output folder: {tmp_path}\n")
    sim = ref.Simulation(pf, num_threads=1)
    f0 = sim.fields()
    for loop in range(6):
        sim.iteration(loop, 20000)
    f1 = sim.fields()
    sim.close()
    cells = np.ascontiguousarray(np.stack([f0[0], f0[2], f0[3], f0[1]], 1))
    xmetal = np.ascontiguousarray(f0[4:16].T)
    anchor, sides, ncell = np.array([-3 * PC] * 3), np.array([6 * PC] * 3), np.array([nc] * 3, np.int32)
    src = np.array([0.1 * PC, 0.2 * PC, -0.1 * PC])
    abund = np.array([0.1, 2.2e-4, 4e-5, 3.3e-4, 5e-5, 9e-6])
    tpar = np.array([0., 0., 0.75, ref.convert(1.33333, "kpc", "m"), 4000., 1e-3, 100.])
    hostcheck.hc_simulation(p(anchor), p(sides), p(ncell), p(src), C.c_double(1e49), C.c_double(40000.), p(abund), p(tpar),
                            C.c_int(1), C.c_uint32(6), C.c_uint64(20000), C.c_int(321), p(cells), p(xmetal))
    assert np.array_equal(cells[:, 3], f1[1])                  # temperature
    assert np.array_equal(cells[:, 1], f1[2]) and np.array_equal(cells[:, 2], f1[3])   # x_H, x_He
    assert np.array_equal(xmetal.T, f1[4:16])                  # the 12 metal fractions
    assert (f1[1] > 4000.).mean() > 0.3 and np.unique(f1[1]).size > 100     # the temperature solve really ran
