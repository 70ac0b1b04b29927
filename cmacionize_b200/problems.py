"""The benchmark problems of BASELINE.json expressed as C-ABI calls.

Each builder mirrors one of the reference's parameter files
(/root/reference/benchmarks/*.param, values quoted in SURVEY.md §8d) and returns
a configured :class:`Context` plus the host-side initial cell arrays.  Unit
conversions reproduce UnitConverter (src/UnitConverter.hpp:103-170): pc =
3.086e16 m, cm^-3 = 1e6 m^-3, eV -> Hz = value*eV*(1/h).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import capi
from .capi import Context

PC = 3.086e16
EV = 1.6021766208e-19
PLANCK = 6.626070040e-34


def ev_to_hz(ev: float) -> float:
    return (ev * EV) * (1. / PLANCK)


@dataclass
class Problem:
    name: str
    ctx: Context
    number_density: np.ndarray
    temperature: np.ndarray
    ionic_fractions: np.ndarray
    n_packets: int
    n_iterations: int
    seed: int = 42
    steps_per_packet: float = 0.   # measured by the oracle for the roofline
    bytes_per_step: float = 0.     # algorithmic bytes per packet-cell crossing
    meta: dict = field(default_factory=dict)

    def upload(self):
        self.ctx.upload_cells(self.number_density, self.temperature, self.ionic_fractions)


def cell_midpoints(anchor, sides, ncell):
    """CartesianDensityGrid::get_cell_midpoint (CartesianDensityGrid.hpp:85-89)"""
    mids = []
    for d in range(3):
        cs = sides[d] / ncell[d]
        lo = anchor[d] + cs * np.arange(ncell[d], dtype=np.float64)
        mids.append(lo + 0.5 * cs)
    X, Y, Z = np.meshgrid(*mids, indexing="ij")
    return X.reshape(-1), Y.reshape(-1), Z.reshape(-1)


def initial_fractions(ncells: int, xH: float = 1.e-6) -> np.ndarray:
    # HomogeneousDensityFunction.hpp:99-108 / DensityValues.hpp:65-71: x_H from the
    # parameter file (default 1e-6), x_He = 1e-6, metals 0
    x = np.zeros((capi.NUM_IONS, ncells))
    x[0] = xH
    x[1] = 1.e-6
    return x


def stromgren(ncell: int = 64, n_packets: int = 1_000_000, n_iterations: int = 20,
              diffuse: bool = False, device: int = 0) -> Problem:
    """benchmarks/stromgren.param (+ stromgren_diffuse.param with diffuse=True)"""
    anchor = [-5. * PC] * 3
    sides = [10. * PC] * 3
    ctx = Context(anchor, sides, [ncell] * 3, device=device)
    nc = ctx.ncells
    ctx.set_abundances()  # AbundanceModel default: all 0
    sig = np.zeros(capi.NUM_IONS)
    sig[0] = 6.3e-18 * 1.e-4  # cm^2 -> m^2
    ctx.set_cross_sections(capi.CROSS_SECTIONS_FIXED_VALUE, sig)
    rr = np.zeros(capi.NUM_IONS)
    rr[0] = 4.e-13 * 1.e-6  # cm^3 s^-1 -> m^3 s^-1
    ctx.set_recombination_rates(capi.RECOMBINATION_FIXED_VALUE, rr)
    ctx.set_sources([[0., 0., 0.]], [1.], 4.26e49)
    ctx.set_spectrum(capi.SPECTRUM_MONOCHROMATIC, ev_to_hz(13.6))
    ctx.set_reemission(capi.REEMISSION_PHYSICAL if diffuse else capi.REEMISSION_NONE)
    ctx.set_temperature_params(do_temperature_calculation=False)
    prob = Problem("stromgren_diffuse" if diffuse else "stromgren", ctx,
                   np.full(nc, 100. * 1.e6), np.full(nc, 8000.), initial_fractions(nc), n_packets,
                   n_iterations, bytes_per_step=24.)
    prob.upload()
    return prob


def lexington(which: int = 20, ncell: int = 64, n_packets: int = 100_000_000,
              n_iterations: int = 20, device: int = 0) -> Problem:
    """benchmarks/lexingtonHII20.param / lexingtonHII40.param with their .yml blocks"""
    if which == 20:
        half, T_star, Q = 3. * PC, 20000., 1.e49
    elif which == 40:
        half, T_star, Q = 5. * PC, 40000., 4.26e49
    else:
        raise ValueError("Lexington benchmark is HII20 or HII40")
    anchor = [-half] * 3
    sides = [2. * half] * 3
    ctx = Context(anchor, sides, [ncell] * 3, device=device)
    nc = ctx.ncells
    ctx.set_abundances(He=0.1, C_=2.2e-4, N=4.e-5, O=3.3e-4, Ne=5.e-5, S=9.e-6)
    ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
    ctx.set_recombination_rates(capi.RECOMBINATION_VERNER)
    ctx.set_sources([[0., 0., 0.]], [1.], Q)
    ctx.set_spectrum(capi.SPECTRUM_PLANCK, T_star)
    ctx.set_reemission(capi.REEMISSION_PHYSICAL)
    ctx.set_temperature_params(do_temperature_calculation=True, pah_heating_factor=0.)
    # BlockSyntaxDensityFunction (BlockSyntaxDensityFunction.hpp:151-199): later blocks win.
    # block[0] cube 100 cm^-3 8000 K; block[1] sphere of diameter 6e18 cm: vacuum, 0 K
    X, Y, Z = cell_midpoints(anchor, sides, [ncell] * 3)
    n = np.full(nc, 100. * 1.e6)
    T = np.full(nc, 8000.)
    s = 6.e18 * 0.01
    # BlockSyntaxBlock::is_inside (BlockSyntaxBlock.hpp:91-106), exponent 2
    r = np.zeros(nc)
    for c in (X, Y, Z):
        x = 2. * np.abs(c - 0.) / s
        r += np.power(x, 2.)
    r = np.power(r, 1. / 2.)
    inside = r <= 1.
    n[inside] = 0.
    T[inside] = 0.
    prob = Problem(f"lexingtonHII{which}", ctx, n, T, initial_fractions(nc), n_packets,
                   n_iterations, bytes_per_step=152.)
    prob.upload()
    return prob


def synthetic_clumpy(ncell: int = 256, n_sources: int = 16, n_packets: int = 1_000_000_000,
                     n_iterations: int = 10, variant: str = "H", device: int = 0) -> Problem:
    """log-normal clumpy density, i.i.d. per 4^3 block,
    sigma_ln = 1, seed 1234; S sources uniformly placed (seed 4321), equal weights,
    Q_tot = S x 4.26e49 s^-1; variant "H": monochromatic + FixedValue (24 B/step class),
    variant "Lexington": Planck 40000 K + Verner + metals + T solve (152 B/step class)."""
    anchor = [-5. * PC] * 3
    sides = [10. * PC] * 3
    ctx = Context(anchor, sides, [ncell] * 3, device=device)
    nc = ctx.ncells
    rng = np.random.default_rng(1234)
    nb = max(ncell // 4, 1)
    g = rng.standard_normal((nb, nb, nb))
    rep = ncell // nb
    g = np.repeat(np.repeat(np.repeat(g, rep, 0), rep, 1), rep, 2).reshape(-1)
    sigma_ln = 1.
    n = 100. * 1.e6 * np.exp(sigma_ln * g - 0.5 * sigma_ln * sigma_ln)
    rs = np.random.default_rng(4321)
    pos = (np.array(anchor) + np.array(sides) * rs.uniform(0.1, 0.9, (n_sources, 3)))
    ctx.set_sources(pos, np.full(n_sources, 1. / n_sources), n_sources * 4.26e49)
    if variant == "H":
        ctx.set_abundances()
        sig = np.zeros(capi.NUM_IONS); sig[0] = 6.3e-22
        ctx.set_cross_sections(capi.CROSS_SECTIONS_FIXED_VALUE, sig)
        rr = np.zeros(capi.NUM_IONS); rr[0] = 4.e-19
        ctx.set_recombination_rates(capi.RECOMBINATION_FIXED_VALUE, rr)
        ctx.set_spectrum(capi.SPECTRUM_MONOCHROMATIC, ev_to_hz(13.6))
        ctx.set_reemission(capi.REEMISSION_NONE)
        ctx.set_temperature_params(do_temperature_calculation=False)
        bps = 24.
    else:
        ctx.set_abundances(He=0.1, C_=2.2e-4, N=4.e-5, O=3.3e-4, Ne=5.e-5, S=9.e-6)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_recombination_rates(capi.RECOMBINATION_VERNER)
        ctx.set_spectrum(capi.SPECTRUM_PLANCK, 40000.)
        ctx.set_reemission(capi.REEMISSION_PHYSICAL)
        ctx.set_temperature_params(do_temperature_calculation=True)
        bps = 152.
    prob = Problem(f"synthetic_clumpy_{ncell}_{variant}", ctx, n, np.full(nc, 8000.),
                   initial_fractions(nc), n_packets, n_iterations, seed=1234, bytes_per_step=bps,
                   meta=dict(n_sources=n_sources))
    prob.upload()
    return prob


def run_iteration(prob: Problem, loop: int, n_packets: int | None = None, packet_offset: int = 0,
                  allreduce=None, want_counters: bool = False):
    """One pass of IonizationSimulation::run's loop body (IonizationSimulation.cpp:359-643):
    reset_grid -> set_reemission_probabilities -> shoot -> [all-reduce] -> state update."""
    ctx = prob.ctx
    ctx.reset_accumulators()
    ctx.update_reemission_probabilities()
    out = ctx.shoot(prob.n_packets if n_packets is None else n_packets, packet_offset=packet_offset,
                    seed=prob.seed, iteration=loop, want_counters=want_counters)
    if allreduce is not None:
        allreduce(ctx)
    ctx.update_state(loop, 0.)
    return out


def run(prob: Problem, n_iterations: int | None = None, n_packets: int | None = None):
    for loop in range(prob.n_iterations if n_iterations is None else n_iterations):
        run_iteration(prob, loop, n_packets)
    prob.ctx.synchronize()
