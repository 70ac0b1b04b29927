"""GPU tier, BASELINE.json configs[4] and the north-star grid: parity at 256^3 (16.8 M cells), where the walk
uses what the small grids never reach — 24-bit cell indices, the planar H-only accumulators, the coherent march
(march_lean_kernel: ordered queue, wall tables, in-warp sums, hot-cell replicas of 16 sources), rounds of up to
64 Mi packets — against the oracle (the reference's CartesianDensityGrid::interact on a 256^3 grid of its own
cells, ~5 GB of host memory) on identical packets."""
import os

import numpy as np
import pytest

from cases import march_case

pytestmark = pytest.mark.gpu

PC = 3.086e16


def test_traversal_bit_exact_on_a_256_cubed_grid(cmib, ref):
    """1e5 explicit packets through 256^3 random cells: same cells in the same order, same end cell, bit-identical
    end positions, accumulators to 1e-12 (the contract of tests/test_gpu_march.py at the full grid size)."""
    c = march_case("clumpy256", 100000)
    mt = 512   # the longest walks cross ~3 x 256 cells; the trace keeps the first 512 of each
    r = ref.interact(c["anchor"], c["sides"], c["ncell"], c["periodic"], c["n"], c["xH"], c["xHe"], c["pos"], c["dir"],
                     c["sigma"], c["sigma_He_corr"], c["nu"], c["weight"], c["tau"], max_trace=mt)
    with cmib.Context(c["anchor"], c["sides"], c["ncell"], c["periodic"]) as ctx:
        nc = ctx.ncells
        x = np.zeros((14, nc)); x[0] = c["xH"]; x[1] = c["xHe"]
        ctx.upload_cells(c["n"], np.full(nc, 8000.), x)
        del x
        ctx.reset_accumulators()
        fpos, fcell, nsteps, trace = ctx.march_packets(c["pos"], c["dir"], c["sigma"], c["sigma_He_corr"], c["nu"],
                                                       c["weight"], c["tau"], max_trace=mt)
        J, heat = ctx.download_accumulators()
    assert nsteps.max() > 256 and (fcell >= 2 ** 23).any()           # long walks, large cell indices
    assert np.array_equal(nsteps, r["nsteps"])
    assert np.array_equal(trace, r["trace"])
    assert np.array_equal(fcell, r["final_cell"])
    assert np.array_equal(fpos, r["final_pos"])
    scale = np.abs(r["J"]).max(axis=1, keepdims=True)
    assert (np.abs(J - r["J"]) <= 1e-12 * np.maximum(scale, 1e-300)).all()
    hs = np.abs(r["heat"]).max(axis=1, keepdims=True)
    assert (np.abs(heat - r["heat"]) <= 1e-12 * np.maximum(hs, 1e-300)).all()
    assert np.array_equal(J == 0., r["J"] == 0.)


@pytest.mark.parametrize("order", ["coherent", "default", "emission"])
def test_production_shoot_on_clumpy_256_equals_oracle(cmib, ref, order):
    """The production shoot on the synthetic clumpy 256^3 grid with 16 sources (H-only planes, hot-cell replicas
    of 16 sources, the coherent march forced / chosen by the library's rule / emission order): the packets it
    draws, exported and pushed through the oracle's interact(), give its accumulators to 1e-11, its packet fates
    and its number of cell crossings."""
    from cmacionize_b200 import problems
    npk = 1_200_000     # >= 2^20: the size from which the rule picks the coherent march on this grid
    prob = problems.synthetic_clumpy(ncell=256, n_packets=npk)
    ctx = prob.ctx
    rng = np.random.default_rng(3)
    x = prob.ionic_fractions.copy()
    x[0] = np.exp(rng.uniform(np.log(1e-6), np.log(1e-3), ctx.ncells))   # mostly ionised: long walks
    ctx.upload_cells(prob.number_density, prob.temperature, x)
    env = {"coherent": "2", "default": None, "emission": "0"}[order]
    if env is None:
        os.environ.pop("CMIB_SORT", None)
    else:
        os.environ["CMIB_SORT"] = env
    try:
        ctx.reset_accumulators()
        tw, tc = ctx.shoot(npk, seed=1234, iteration=2)
    finally:
        os.environ.pop("CMIB_SORT", None)
    J, heat = ctx.download_accumulators()
    crossings, emissions = ctx.shoot_statistics()
    pk = ctx.sample_packets(npk, seed=1234, iteration=2)
    ctx.close()
    r = ref.interact([-5 * PC] * 3, [10 * PC] * 3, [256] * 3, [0, 0, 0], prob.number_density, x[0], x[1], pk["pos"],
                     pk["dir"], pk["sigma"], pk["sigma_He_corr"], pk["nu"], np.ones(npk), pk["tau"])
    scale = r["J"][0].max()
    assert np.abs(J[0] - r["J"][0]).max() <= 1e-11 * scale
    assert np.array_equal(J[0] == 0., r["J"][0] == 0.)
    assert np.array_equal(J[1:], np.zeros_like(J[1:])) and np.array_equal(heat, np.zeros_like(heat))
    absorbed = (r["final_cell"] >= 0).sum()
    assert tw == npk and tc[3] == absorbed and tc[0] == npk - absorbed and emissions == npk
    assert crossings == r["nsteps"].sum() or abs(crossings - r["nsteps"].sum()) <= npk  # vacuum-free grid: one per cell


@pytest.mark.parametrize("grid", ["stromgren256", "clumpy256"])
def test_full_size_checksum_of_the_accumulation_at_256_cubed(cmib, grid):
    """1.6e7 packets (a full round of the coherent march) on the 256^3 grids in their ionised steady state:
    sum_cells n x_H J_H = optical depth traversed (summed per packet inside the walk) to 1e-9, packet
    conservation exact.  A lost, doubled or mis-addressed in-warp sum breaks it (cells differ in n and x)."""
    from cmacionize_b200 import problems
    npk = 16_000_000
    prob = problems.stromgren(ncell=256, n_packets=npk) if grid == "stromgren256" else problems.synthetic_clumpy(ncell=256, n_packets=npk)
    ctx = prob.ctx
    for loop in range(4):
        problems.run_iteration(prob, loop, n_packets=4_000_000)
    n, T, x, _ = ctx.download_cells()
    rng = np.random.default_rng(1)
    n = n * rng.uniform(0.5, 1.5, n.size)
    ctx.upload_cells(n, T, x)
    ctx.reset_accumulators()
    tw, tc = ctx.shoot(npk, seed=11, iteration=4)
    crossings, emissions = ctx.shoot_statistics()
    tau = ctx.shoot_optical_depth()
    J, heat = ctx.download_accumulators()
    ctx.close()
    assert tw == npk and tc.sum() == npk and emissions == npk and crossings > 50 * npk
    lhs = float(np.sum(n * x[0] * J[0]))
    assert abs(lhs - tau) <= 1e-9 * tau, (lhs, tau)
