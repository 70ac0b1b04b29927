"""Pin the oracle: the compiled reference (oracle/_ref) must reproduce every
known-answer fixture the reference's own unit tests hold for this path, at the
reference's own tolerances.  `assert_values_equal_rel(a,b,t)`
in the reference means |a-b| <= t*|a+b| (test/Assert.hpp:36-73)."""
import numpy as np

from cases import ABUNDANCES

EV = 1.6021766208e-19
H = 6.626070040e-34


def assert_rel(a, b, tol):
    a = np.asarray(a, float)
    b = np.asarray(b, float)
    bad = np.abs(a - b) > tol * np.abs(a + b)
    assert not bad.any(), f"{bad.sum()} values differ, worst {np.max(np.abs(a-b)/np.maximum(np.abs(a+b),1e-300))}"


def test_verner_cross_sections(ref, golden):
    # test/testVernerCrossSections.cpp:42-170
    g = golden["verner_xsec"]
    nu = (g[:, 0] * 13.6 * EV) * (1. / H)
    sigma = ref.verner_cross_sections(nu)
    assert_rel(sigma * 1e4 * 1e18, g[:, 1:], 1e-9)


def test_verner_recombination_rates(ref, golden):
    # test/testVernerRecombinationRates.cpp:40-153
    g = golden["verner_rec"]
    alpha = ref.verner_recombination_rates(g[:, 0])
    assert_rel(alpha * 1e6, g[:, 1:], 1e-14)


def test_charge_transfer(ref, golden):
    # test/testChargeTransferRates.cpp:44-150
    g = golden["kingdon_ferland"]
    ion = {(6, 4): 3, (7, 1): 4, (7, 2): 4, (7, 3): 5, (7, 4): 6, (8, 1): 7, (8, 2): 7, (8, 3): 8,
           (10, 3): 10, (16, 3): 11, (16, 4): 12, (16, 5): 13}
    ct = ref.charge_transfer(g[:, 2] * 1e-4)
    n_checked = 0
    for r, row in enumerate(g):
        stage, atom = int(row[0]), int(row[1])
        if (atom, stage) not in ion:
            continue
        k = ion[(atom, stage)]
        if stage > 1:
            assert_rel(ct[r, 0, k] * 1e6, row[3], 1e-6)
            n_checked += 1
        if (atom, stage) in ((7, 1), (8, 1)):
            assert_rel(ct[r, 1, k] * 1e6, row[4], 1e-6)
            n_checked += 1
    assert n_checked > 1000


def test_line_cooling(ref, golden):
    # test/testLineCoolingData.cpp:127-149
    g = golden["linecool"]
    cool = ref.linecooling_get_cooling(g[:, 0], g[:, 1] * 1e6, g[:, 2:15])
    assert_rel(cool * 1e7, g[:, 15], 1e-6)


def test_solve5_random_systems(ref):
    # test/testLineCoolingData.cpp:87-124
    rng = np.random.default_rng(42)
    A = rng.uniform(0, 1, (10000, 5, 5))
    for i in range(5):
        A[:, i, i] = 1.
    B = rng.uniform(0, 1, (10000, 5))
    _, X, st = ref.solve5(A.reshape(-1, 25), B)
    assert (st == 0).all()
    back = np.einsum("nij,nj->ni", A, X)
    assert np.allclose(back, B, rtol=1e-11, atol=1e-11)


def test_ionization_state(ref, golden):
    # test/testIonizationStateCalculator.cpp:46-222 (abundances 0.1,0,0,0,0,0)
    g = golden["h0"]
    J = np.ascontiguousarray(g[:, :14].T)
    x, _ = ref.ionization_state(1., 1., [0.1, 0, 0, 0, 0, 0], 1, None, J, np.zeros((2, len(g))),
                                g[:, 15] * 1e6, g[:, 14])
    assert_rel(x.T, g[:, 16:30], 1e-9)


def test_cooling_heating_balance(ref, golden):
    # test/testTemperatureCalculator.cpp:97-178
    g = golden["ioneng"]
    h0, he0, gain, loss, metals = ref.cooling_heating_balance(
        g[:, 16], g[:, 19] * 1e6, np.ascontiguousarray(g[:, :14]),
        np.ascontiguousarray(g[:, 14:16]) * 1e-7, ABUNDANCES, 1., 0., 0.75)
    assert_rel(h0, g[:, 20], 1e-6)
    assert_rel(he0, g[:, 21], 1e-6)
    assert_rel(gain, g[:, 17] * 0.1 * 1e-20, 1e-6)
    assert_rel(loss, g[:, 18] * 0.1 * 1e-20, 1e-6)
    assert_rel(metals, g[:, 22:34], 1e-6)


def test_calculate_temperature(ref, golden):
    # test/testTemperatureCalculator.cpp:179-322
    g = golden["tbal"]
    g = g[g[:, 16] <= 30000.]
    T, x, _ = ref.temperature(1., 1., ABUNDANCES, np.ascontiguousarray(g[:, :14].T),
                              np.ascontiguousarray(g[:, 14:16].T) * 1e-7, g[:, 17] * 1e6, g[:, 16],
                              pahfac=1., crfac=0., crlim=1., crscale=0.)
    assert_rel(T, np.minimum(30000., g[:, 32]), 1e-4)
    assert_rel(x[0], np.minimum(1., g[:, 18]), 1e-4)
    assert_rel(x[1:].T, g[:, 19:32], 1e-4)


def test_reemission_probabilities(ref, golden):
    # test/testPhysicalDiffuseReemissionHandler.cpp:40-76
    g = golden["probset"]
    p = ref.reemission_probabilities(g[:, 0])
    assert_rel(p, g[:, 1:6], 1e-15)


def test_cartesian_grid_geometry(ref):
    # test/testCartesianDensityGrid.cpp:65-68, 467-475: cell of (0.51,0.51,0.51) in a 16^3
    # unit box is 8*256+8*16+8; a packet shot through an almost transparent box leaves it
    nc = 16 ** 3
    out = ref.interact([0, 0, 0], [1, 1, 1], [16, 16, 16], [0, 0, 0], np.full(nc, 1.), np.full(nc, 1e-6),
                       np.zeros(nc), [[0.51, 0.51, 0.51]], [[1., 0., 0.]], np.zeros((1, 14)) + 1e-30,
                       [0.], [3.3e15], [1.], [1e-40], max_trace=4)
    assert out["trace"][0, 0] == 8 * 256 + 8 * 16 + 8
    assert out["final_cell"][0] == 8 * 256 + 8 * 16 + 8  # tau exhausted in the first cell
    out = ref.interact([0, 0, 0], [1, 1, 1], [16, 16, 16], [0, 0, 0], np.full(nc, 1.), np.full(nc, 1e-6),
                       np.zeros(nc), [[0.5, 0.5, 0.5]], [[1., 1., 1.] / np.sqrt(3.)],
                       np.zeros((1, 14)) + 1e-30, [0.], [3.3e15], [1.], [1.], max_trace=16)
    assert out["final_cell"][0] == -1
    # a body-diagonal ray from a cell corner crosses corners: x, y and z step together
    assert list(out["trace"][0, :8]) == [(8 + k) * 256 + (8 + k) * 16 + (8 + k) for k in range(8)]
