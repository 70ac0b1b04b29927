"""GPU tier: every physics function of the hot path, evaluated ON THE DEVICE through
the C ABI (cmib_eval_*), against the compiled reference on the same inputs and
against the reference's golden vectors.

Tolerances: the library is compiled with -fmad=false (no FMA contraction, like the
reference), so the only arithmetic difference is CUDA's libm (pow/exp/log/cbrt within
1-2 ulp of glibc's).  Plain fits agree to ~1e-14; functions that amplify an ulp
(exp of a large argument in the dielectronic fits, the ill-conditioned 5x5 level
population systems, the cancelling closed form 1 + a(1 - sqrt(1 + 2/a)), the secant
temperature iteration) agree to the bounds stated at each assert, which leave about one
order of magnitude of head room over what was measured on B200 (profiles/parity_r01.md).
"""
import json
import os
from pathlib import Path

import numpy as np
import pytest

from cases import ABUNDANCES, state_cells
from conftest import rel_err

pytestmark = pytest.mark.gpu

EV = 1.6021766208e-19
H = 6.626070040e-34
MEASURED = {}


@pytest.fixture(scope="module")
def ctx(cmib):
    c = cmib.Context([0, 0, 0], [1, 1, 1], [4, 4, 4])
    c.set_abundances(*ABUNDANCES)
    c.set_cross_sections(cmib.capi.CROSS_SECTIONS_VERNER)
    c.set_recombination_rates(cmib.capi.RECOMBINATION_VERNER)
    yield c
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_physics.json").write_text(json.dumps(MEASURED, indent=1))
    c.close()


def _flush():
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_physics.json").write_text(json.dumps(MEASURED, indent=1))


def record(name, value):
    MEASURED[name] = float(value)
    _flush()
    return value


def golden_rel(a, b, tol):
    a = np.asarray(a); b = np.asarray(b)
    assert (np.abs(a - b) <= tol * np.abs(a + b)).all()


def test_cross_sections(ctx, ref, golden):
    rng = np.random.default_rng(1)
    nu = 3.288e15 * np.exp(rng.uniform(-0.1, np.log(6.), 50000))
    got = ctx.eval_cross_sections(nu)
    want = ref.verner_cross_sections(nu)
    assert np.array_equal(got == 0., want == 0.)  # identical thresholds
    # device: x^a z^b evaluated as exp(a ln x + b ln z) (cross_sections.cuh) -> ~1e-14, not 1e-16
    assert record("xsec_vs_oracle", rel_err(got, want)) < 5e-13
    g = golden["verner_xsec"]
    nu = (g[:, 0] * 13.6 * EV) * (1. / H)
    golden_rel(ctx.eval_cross_sections(nu) * 1e22, g[:, 1:], 1e-9)


def test_fixed_value_cross_sections(cmib):
    with cmib.Context([0, 0, 0], [1, 1, 1], [2, 2, 2]) as c:
        f = np.arange(1, 15) * 1e-22
        c.set_cross_sections(cmib.capi.CROSS_SECTIONS_FIXED_VALUE, f)
        got = c.eval_cross_sections([1e15, 3.3e15, 1e17])
        assert np.array_equal(got, np.tile(f, (3, 1)))


def test_recombination_rates(ctx, ref, golden):
    rng = np.random.default_rng(2)
    T = np.exp(rng.uniform(np.log(50.), np.log(1e7), 50000))
    assert record("rec_vs_oracle", rel_err(ctx.eval_recombination_rates(T),
                                           ref.verner_recombination_rates(T))) < 1e-10  # exp(-T0/T), |arg| up to 1e4
    g = golden["verner_rec"]
    golden_rel(ctx.eval_recombination_rates(g[:, 0]) * 1e6, g[:, 1:], 1e-13)


def test_charge_transfer(ctx, ref):
    rng = np.random.default_rng(3)
    T4 = np.exp(rng.uniform(np.log(1e-4), np.log(1e3), 50000))
    assert record("ct_vs_oracle", rel_err(ctx.eval_charge_transfer(T4), ref.charge_transfer(T4))) < 1e-13


def test_reemission_probabilities(ctx, ref, golden):
    g = golden["probset"]
    golden_rel(ctx.eval_reemission_probabilities(g[:, 0]), g[:, 1:6], 1e-14)
    T = np.linspace(500., 40000., 10000)
    assert record("reemit_prob_vs_oracle", rel_err(ctx.eval_reemission_probabilities(T),
                                                   ref.reemission_probabilities(T))) < 1e-13


def test_line_cooling(ctx, ref, golden):
    g = golden["linecool"]
    golden_rel(ctx.eval_line_cooling(g[:, 0], g[:, 1] * 1e6, g[:, 2:15]) * 1e7, g[:, 15], 1e-6)
    rng = np.random.default_rng(4)
    n = 20000
    T = np.exp(rng.uniform(np.log(3000.), np.log(40000.), n))
    ne = np.exp(rng.uniform(np.log(1e3), np.log(1e12), n))
    ne[:10] = 0.
    ab = rng.uniform(0, 1e-4, (n, 13))
    got = ctx.eval_line_cooling(T, ne, ab)
    assert (got[:10] == 1e-99).all()
    assert record("linecool_vs_oracle", rel_err(got, ref.linecooling_get_cooling(T, ne, ab))) < 1e-6  # the reference's own get_cooling tolerance


def test_solve5(ctx, ref):
    rng = np.random.default_rng(5)
    A = rng.uniform(-1, 1, (20000, 25))
    B = rng.uniform(-1, 1, (20000, 5))
    A[:5] = 0.
    _, X, st = ctx.eval_solve5(A, B)
    _, Xr, sr = ref.solve5(A, B)
    assert np.array_equal(st, sr) and (st[:5] == 1).all()
    # same pivots, same operation order; only FMA contraction differs
    err = np.abs(X[5:] - Xr[5:]) / np.maximum(np.abs(Xr[5:]).max(axis=1, keepdims=True), 1e-300)
    assert record("solve5_vs_oracle", err.max()) < 1e-9
    back = np.einsum("nij,nj->ni", A[5:].reshape(-1, 5, 5), X[5:])
    assert np.allclose(back, B[5:], rtol=1e-9, atol=1e-9)


def test_ionization_state(ctx, ref, golden, cmib):
    g = golden["h0"]
    with cmib.Context([0, 0, 0], [1, 1, 1], [2, 2, 2]) as c:
        c.set_abundances(He=0.1)
        c.set_recombination_rates(cmib.capi.RECOMBINATION_VERNER)
        x, _ = c.eval_ionization_state(1., 1., np.ascontiguousarray(g[:, :14].T), np.zeros((2, len(g))),
                                       g[:, 15] * 1e6, g[:, 14])
        golden_rel(x.T, g[:, 16:30], 1e-9)
    J, heat, nd, T = state_cells(golden, reps=40)
    x, ho = ctx.eval_ionization_state(1.3, 1.3 * H, J, heat, nd, T)
    xr, hr = ref.ionization_state(1.3, 1.3 * H, ABUNDANCES, 1, None, J, heat, nd, T)
    assert record("ionstate_vs_oracle", rel_err(x, xr)) < 1e-9
    assert rel_err(ho, hr) < 1e-14


def test_hydrogen_only_fixed_rates(cmib, ref, golden):
    """Stroemgren configuration: A_He = 0 -> closed form; FixedValue alpha with zeros for
    the metals gives 0/0 = NaN metal fractions in the reference — reproduced, not hidden."""
    J, heat, nd, T = state_cells(golden, reps=5)
    fixed = np.array([4e-19] + [0.] * 13)
    with cmib.Context([0, 0, 0], [1, 1, 1], [2, 2, 2]) as c:
        c.set_abundances()
        c.set_recombination_rates(cmib.capi.RECOMBINATION_FIXED_VALUE, fixed)
        x, _ = c.eval_ionization_state(2., 2. * H, J, heat, nd, T)
    xr, _ = ref.ionization_state(2., 2. * H, np.zeros(6), 0, fixed, J, heat, nd, T)
    # x = 1 + a(1 - sqrt(1 + 2/a)) cancels ~log10(a) digits: compare on the ionized fraction
    # 1 - x (well conditioned) tightly and on x itself to the cancellation-limited bound
    assert record("h_only_ionized_vs_oracle", rel_err(1. - x[0], 1. - xr[0])) < 1e-12
    assert record("h_only_vs_oracle", rel_err(x[0], xr[0])) < 1e-6
    assert np.array_equal(np.isnan(x), np.isnan(xr))


def test_cooling_heating_balance(ctx, ref, golden):
    g = golden["ioneng"]
    ctx.set_temperature_params(do_temperature_calculation=True, pah_heating_factor=1.,
                               cosmic_ray_heating_factor=0., cosmic_ray_heating_scale_length=0.75)
    T = g[:, 16]; nd = g[:, 19] * 1e6
    j = np.ascontiguousarray(g[:, :14]); h = np.ascontiguousarray(g[:, 14:16]) * 1e-7
    h0, he0, gain, loss, metals = ctx.eval_cooling_heating_balance(T, nd, j, h)
    golden_rel(h0, g[:, 20], 1e-6)
    golden_rel(he0, g[:, 21], 1e-6)
    golden_rel(gain, g[:, 17] * 0.1 * 1e-20, 1e-6)
    golden_rel(loss, g[:, 18] * 0.1 * 1e-20, 1e-6)
    golden_rel(metals, g[:, 22:34], 1e-6)
    r = ref.cooling_heating_balance(T, nd, j, h, ABUNDANCES, 1., 0., 0.75)
    worst = max(rel_err(a, b) for a, b in zip((h0, he0, gain, loss, metals), r))
    assert record("balance_vs_oracle", worst) < 1e-5  # golden tolerance of the reference: 1e-6 .. 1e-4


def test_calculate_temperature(ctx, ref, golden):
    g = golden["tbal"]
    g = g[g[:, 16] <= 30000.]
    ctx.set_temperature_params(do_temperature_calculation=True, pah_heating_factor=1.,
                               cosmic_ray_heating_factor=0., cosmic_ray_heating_limit=1.,
                               cosmic_ray_heating_scale_length=0.)
    T, x, _ = ctx.eval_temperature(1., 1., np.ascontiguousarray(g[:, :14].T),
                                   np.ascontiguousarray(g[:, 14:16].T) * 1e-7, g[:, 17] * 1e6, g[:, 16])
    golden_rel(T, np.minimum(30000., g[:, 32]), 1e-4)
    golden_rel(x[0], np.minimum(1., g[:, 18]), 1e-4)
    golden_rel(x[1:].T, g[:, 19:32], 1e-4)
    # wide randomised sweep incl. cosmic rays, J = 0, vacuum, T clamps
    J, heat, nd, T0 = state_cells(golden, reps=40)
    n = nd.size
    rng = np.random.default_rng(11)
    crf = rng.uniform(-1, 2, n)
    mz = rng.uniform(-1e19, 1e19, n)
    ctx.set_temperature_params(do_temperature_calculation=True, pah_heating_factor=0.5,
                               cosmic_ray_heating_factor=0.2, cosmic_ray_heating_limit=0.75,
                               cosmic_ray_heating_scale_length=1e19)
    Tg, xg, hg = ctx.eval_temperature(1., 1., J, heat, nd, T0, crf, mz)
    Tr, xr, hr = ref.temperature(1., 1., ABUNDANCES, J, heat, nd, T0, pahfac=0.5, crfac=0.2,
                                 crlim=0.75, crscale=1e19, cr_factor=crf, midz=mz)
    # the secant iteration stops on |gain-loss| <= 1e-3 gain: a 1-ulp difference can change
    # the iteration count of a cell sitting on that edge, which moves T by up to ~eps.
    # Require agreement to 1e-6 for >= 95 % of cells and the reference's own
    # convergence tolerance for every cell.
    dT = np.abs(Tg - Tr) / Tr
    record("temperature_median_rel", np.median(dT))
    record("temperature_max_rel", dT.max())
    record("temperature_frac_gt_1e-9", np.mean(dT > 1e-9))
    bad = np.where(dT > 2e-3)[0]
    MEASURED["temperature_outliers"] = [dict(i=int(i), Tg=float(Tg[i]), Tr=float(Tr[i])) for i in bad[:10]]
    _flush()
    assert np.mean(dT > 1e-6) < 0.05
    # A cell may only be off by more than the solver's tolerance if it is ILL-CONDITIONED IN THE
    # REFERENCE ITSELF: the reference's answer for it must jump when its input temperature is
    # moved by 1e-13 .. 1e-10 relative (measured on B200: 1 such cell of 1160, a jittered fixture
    # row whose heating is ~0 so that the sign of expgain - exploss is rounding noise and the
    # secant step lands on 1e10 K -> 30000 K in one code and < 4000 K -> 500 K in the other).
    assert bad.size <= 0.003 * n
    for i in bad:
        sl = slice(i, i + 1)
        outs = []
        for f in (1. - 1e-10, 1. - 1e-13, 1., 1. + 1e-13, 1. + 1e-10):
            Tp, _, _ = ref.temperature(1., 1., ABUNDANCES, np.ascontiguousarray(J[:, sl]),
                                       np.ascontiguousarray(heat[:, sl]), nd[sl].copy(), T0[sl] * f,
                                       pahfac=0.5, crfac=0.2, crlim=0.75, crscale=1e19,
                                       cr_factor=crf[sl].copy(), midz=mz[sl].copy())
            outs.append(float(Tp[0]))
        assert max(outs) / min(outs) - 1. > 2e-3, (int(i), outs, float(Tg[i]))
        assert min(outs) * (1 - 2e-3) <= Tg[i] <= max(outs) * (1 + 2e-3)
    ok = dT <= 2e-3
    # cells clipped to the 30000 K ceiling hide what the unclipped T0 was: leave them out of the
    # fraction comparison (their T is compared above)
    same = (dT <= 1e-9) & (Tr < 30000.)
    # fractions live in [0, 1]: 1e-5 relative, floored at 1e-12 absolute (metal fractions of
    # 1e-30 carry no information and amplify ulps without bound)
    dx = np.abs(xg[:, same] - xr[:, same]) / (np.abs(xr[:, same]) + 1e-7)
    w = np.unravel_index(np.argmax(dx), dx.shape)
    MEASURED["temperature_x_worst"] = dict(ion=int(w[0]), xg=float(xg[:, same][w]), xr=float(xr[:, same][w]),
                                           Tg=float(Tg[same][w[1]]), Tr=float(Tr[same][w[1]]))
    record("temperature_x_rel_where_T_same", dx.max())
    assert dx.max() < 1e-5
    assert np.array_equal(Tg[ok] == 500., Tr[ok] == 500.)
    assert rel_err(hg[:, ok], hr[:, ok]) < 1e-14
