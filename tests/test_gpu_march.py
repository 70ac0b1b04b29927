"""GPU tier, tiers 1 and 2 of the north star: on IDENTICAL input packets the
device walk must visit the same cells in the same order as
CartesianDensityGrid::interact (bit-exact integer work), end at the bit-identical
position, and accumulate path lengths / optical depths within 1e-6 relative
(observed ~1e-15: only the summation order of the atomics differs)."""
import numpy as np
import pytest

from cases import MARCH_GRIDS, check_wall_intersection_scenarios, march_case, wall_intersection_scenarios
from conftest import rel_err

pytestmark = pytest.mark.gpu


def run_case(cmib, ref, name, npk, mt=256):
    c = march_case(name, npk)
    r = ref.interact(c["anchor"], c["sides"], c["ncell"], c["periodic"], c["n"], c["xH"], c["xHe"],
                     c["pos"], c["dir"], c["sigma"], c["sigma_He_corr"], c["nu"], c["weight"],
                     c["tau"], max_trace=mt)
    with cmib.Context(c["anchor"], c["sides"], c["ncell"], c["periodic"]) as ctx:
        nc = ctx.ncells
        x = np.zeros((14, nc)); x[0] = c["xH"]; x[1] = c["xHe"]
        ctx.upload_cells(c["n"], np.full(nc, 8000.), x)
        ctx.reset_accumulators()
        fpos, fcell, nsteps, trace = ctx.march_packets(c["pos"], c["dir"], c["sigma"],
                                                       c["sigma_He_corr"], c["nu"], c["weight"],
                                                       c["tau"], max_trace=mt)
        J, heat = ctx.download_accumulators()
    return c, r, fpos, fcell, nsteps, trace, J, heat


@pytest.mark.parametrize("name", list(MARCH_GRIDS))
def test_traversal_bit_exact(cmib, ref, name):
    c, r, fpos, fcell, nsteps, trace, J, heat = run_case(cmib, ref, name, 20000)
    assert np.array_equal(nsteps, r["nsteps"])           # same number of cells
    assert np.array_equal(trace, r["trace"])             # same cells, same order
    assert np.array_equal(fcell, r["final_cell"])        # same absorbing cell / escape
    assert np.array_equal(fpos, r["final_pos"])          # bit-identical end position
    # accumulators: J = sum ds*w*sigma -> path-length parity; heat likewise
    scale = np.abs(r["J"]).max(axis=1, keepdims=True)
    assert (np.abs(J - r["J"]) <= 1e-12 * np.maximum(scale, 1e-300)).all()
    assert rel_err(J[r["J"] > 1e-6 * scale], r["J"][r["J"] > 1e-6 * scale]) < 1e-6
    hs = np.abs(r["heat"]).max(axis=1, keepdims=True)
    assert (np.abs(heat - r["heat"]) <= 1e-12 * np.maximum(hs, 1e-300)).all()
    assert np.array_equal(J == 0., r["J"] == 0.)         # untouched cells stay exactly zero


def test_empty_and_degenerate_inputs(cmib, ref):
    with cmib.Context([0, 0, 0], [1, 1, 1], [4, 4, 4]) as ctx:
        nc = ctx.ncells
        x = np.zeros((14, nc)); x[0] = 1.
        ctx.upload_cells(np.full(nc, 1e20), np.full(nc, 8000.), x)
        # zero packets
        fpos, fcell, nsteps, _ = ctx.march_packets(np.zeros((0, 3)), np.zeros((0, 3)), np.zeros((0, 14)),
                                                   np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0))
        assert fpos.shape == (0, 3)
        # tau = 0: no step is taken, packet stays put and reports "no cell" like interact()
        sig = np.zeros((2, 14)); sig[:, 0] = 6.3e-22
        pos = np.array([[0.3, 0.3, 0.3], [2., 2., 2.]])  # second one starts outside the box
        d = np.array([[1., 0., 0.], [1., 0., 0.]])
        fpos, fcell, nsteps, _ = ctx.march_packets(pos, d, sig, np.zeros(2), np.full(2, 3.3e15),
                                                   np.ones(2), np.array([0., 1.]))
        r = ref.interact([0, 0, 0], [1, 1, 1], [4, 4, 4], [0, 0, 0], np.full(nc, 1e20), x[0], x[1],
                         pos[:1], d[:1], sig[:1], np.zeros(1), np.full(1, 3.3e15), np.ones(1),
                         np.array([0.]))
        assert fcell[0] == r["final_cell"][0] and np.array_equal(fpos[0], r["final_pos"][0])
        assert nsteps[0] == 0 and nsteps[1] == 0 and fcell[1] == -1


def test_split_batches_equal_one_batch(cmib):
    """Linearity / sharding property used by the multi-GPU path: marching a batch in two
    halves into the same accumulators gives the same sums as one batch."""
    c = march_case("stromgren64_corner", 20000)
    outs = []
    for parts in (1, 2):
        with cmib.Context(c["anchor"], c["sides"], c["ncell"], c["periodic"]) as ctx:
            nc = ctx.ncells
            x = np.zeros((14, nc)); x[0] = c["xH"]; x[1] = c["xHe"]
            ctx.upload_cells(c["n"], np.full(nc, 8000.), x)
            ctx.reset_accumulators()
            for sl in np.array_split(np.arange(20000), parts):
                ctx.march_packets(c["pos"][sl], c["dir"][sl], c["sigma"][sl], c["sigma_He_corr"][sl],
                                  c["nu"][sl], c["weight"][sl], c["tau"][sl])
            outs.append(ctx.download_accumulators())
    scale = np.abs(outs[0][0]).max()
    assert np.abs(outs[0][0] - outs[1][0]).max() <= 1e-12 * scale


def test_wall_intersection_scenarios_of_the_reference_unit_test(cmib):
    """test/testCartesianDensityGrid.cpp:310-465 through the C ABI (cases.wall_intersection_scenarios)."""
    c = wall_intersection_scenarios()
    with cmib.Context(c["anchor"], c["sides"], c["ncell"], c["periodic"]) as ctx:
        nc = ctx.ncells
        x = np.zeros((14, nc)); x[0] = c["xH"]
        ctx.upload_cells(c["n"], np.full(nc, 8000.), x)
        ctx.reset_accumulators()
        fpos, fcell, nsteps, trace = ctx.march_packets(c["pos"], c["dir"], c["sigma"], c["sigma_He_corr"], c["nu"],
                                                       c["weight"], c["tau"], max_trace=4)
        J, _ = ctx.download_accumulators()
    check_wall_intersection_scenarios(c, fpos, fcell, nsteps, trace, J)


@pytest.mark.parametrize("grid", ["unit16", "stromgren64_corner", "vacuum_holes", "noncubic_periodic_xz"])
def test_integrate_optical_depth_equals_the_reference(cmib, ref, grid):
    """cmib_integrate_optical_depth = DensityGrid::integrate_optical_depth (CartesianDensityGrid.cpp:328-363):
    no libm on this path and every operation separately rounded, so the device result is the reference's
    bit for bit."""
    c = march_case(grid, 4000)
    pos, d = c["pos"], c["dir"]
    if c["periodic"].any():
        keep = np.abs(d[:, 1]) > 0.2
        pos, d = np.ascontiguousarray(pos[keep]), np.ascontiguousarray(d[keep])
    sh = np.ascontiguousarray(c["sigma"][: len(pos), 0])
    she = np.ascontiguousarray(c["sigma_He_corr"][: len(pos)])
    r = ref.integrate_optical_depth(c["anchor"], c["sides"], c["ncell"], c["periodic"], c["n"], c["xH"], c["xHe"],
                                    pos, d, sh, she)
    with cmib.Context(c["anchor"], c["sides"], c["ncell"], c["periodic"]) as ctx:
        nc = ctx.ncells
        x = np.zeros((14, nc)); x[0] = c["xH"]; x[1] = c["xHe"]
        ctx.upload_cells(c["n"], np.full(nc, 8000.), x)
        out = ctx.integrate_optical_depth(pos, d, sh, she)
    assert np.array_equal(out, r)
