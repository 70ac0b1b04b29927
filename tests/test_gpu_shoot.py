"""GPU tier: the fused shoot kernel (emission + walk + accumulation + re-emission).

Self-generated packets cannot be compared bitwise (different RNG by design, and
CUDA's sin/cos/log differ from glibc's in the last ulp — SURVEY.md §7), so:
  * the packets the shoot kernel would draw are exported with cmib_sample_packets
    and pushed through the ORACLE's interact(): shoot's accumulators must equal the
    oracle's on those same packets (this pins emission -> walk -> accumulate
    end to end, at full 1e-6 parity);
  * sampling distributions are checked statistically against the reference's own
    samplers and analytic expectations (the reference's tests do the same,
    test/testPhotonSource.cpp:92-131, test/testPhotonSourceSpectrum.cpp:153-305).
"""
import numpy as np
import pytest

from cases import ABUNDANCES
from conftest import rel_err

pytestmark = pytest.mark.gpu

PC = 3.086e16


def test_shoot_equals_oracle_on_the_same_packets(cmib, ref):
    """No re-emission: shoot == sample_packets -> oracle interact, cell by cell."""
    from cmacionize_b200 import problems
    prob = problems.stromgren(ncell=32, n_packets=20000)
    ctx = prob.ctx
    rng = np.random.default_rng(3)
    n = prob.number_density * np.exp(rng.normal(0, 0.5, ctx.ncells))
    x = prob.ionic_fractions.copy()
    x[0] = np.exp(rng.uniform(np.log(1e-5), np.log(1e-2), ctx.ncells))
    ctx.upload_cells(n, prob.temperature, x)
    ctx.reset_accumulators()
    tw, tc = ctx.shoot(20000, seed=7, iteration=3)
    J, heat = ctx.download_accumulators()
    pk = ctx.sample_packets(20000, seed=7, iteration=3)
    r = ref.interact([-5 * PC] * 3, [10 * PC] * 3, [32] * 3, [0, 0, 0], n, x[0], x[1], pk["pos"],
                     pk["dir"], pk["sigma"], pk["sigma_He_corr"], pk["nu"], np.ones(20000), pk["tau"])
    scale = r["J"][0].max()
    assert np.abs(J[0] - r["J"][0]).max() <= 1e-11 * scale
    assert np.array_equal(J[1:], np.zeros_like(J[1:]))
    assert tw == 20000.
    absorbed = (r["final_cell"] >= 0).sum()
    assert tc[3] == absorbed and tc[0] == 20000 - absorbed
    ctx.close()


def test_full_layout_shoot_equals_oracle(cmib, ref):
    """Planck + Verner (16 accumulators), no re-emission."""
    from cmacionize_b200 import capi
    with cmib.Context([-3 * PC] * 3, [6 * PC] * 3, [24] * 3) as ctx:
        nc = ctx.ncells
        ctx.set_abundances(*ABUNDANCES)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_sources([[0., 0., 0.], [1e16, -2e16, 3e15]], [0.25, 0.75], 1e49)
        ctx.set_spectrum(capi.SPECTRUM_PLANCK, 40000.)
        rng = np.random.default_rng(5)
        n = 1e8 * np.exp(rng.normal(0, 0.5, nc))
        x = np.zeros((14, nc))
        x[0] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), nc))
        x[1] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), nc))
        ctx.upload_cells(n, np.full(nc, 8000.), x)
        ctx.reset_accumulators()
        tw, tc = ctx.shoot(30000, seed=11, iteration=0)
        J, heat = ctx.download_accumulators()
        pk = ctx.sample_packets(30000, seed=11, iteration=0)
        # cross sections of the emitted packets are the oracle's to 5e-13
        assert rel_err(pk["sigma"], ref.verner_cross_sections(pk["nu"])) < 5e-13
        assert np.allclose(pk["sigma_He_corr"], 0.1 * pk["sigma"][:, 1], rtol=1e-15)
        r = ref.interact([-3 * PC] * 3, [6 * PC] * 3, [24] * 3, [0, 0, 0], n, x[0], x[1], pk["pos"],
                         pk["dir"], pk["sigma"], pk["sigma_He_corr"], pk["nu"], np.ones(30000),
                         pk["tau"])
        for k in range(14):
            s = r["J"][k].max()
            assert np.abs(J[k] - r["J"][k]).max() <= 1e-11 * max(s, 1e-300), k
        for k in range(2):
            s = np.abs(r["heat"][k]).max()
            assert np.abs(heat[k] - r["heat"][k]).max() <= 1e-11 * s
        # two sources: 25 % / 75 %
        frac = np.mean(np.all(pk["pos"] == 0., axis=1))
        assert abs(frac - 0.25) < 0.01


def test_emission_statistics(cmib, ref):
    from cmacionize_b200 import capi
    with cmib.Context([-1, -1, -1], [2, 2, 2], [4, 4, 4]) as ctx:
        ctx.set_abundances(*ABUNDANCES)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_sources([[0., 0., 0.]], [1.], 1e49)
        ctx.set_spectrum(capi.SPECTRUM_PLANCK, 40000.)
        n = 1_000_000
        pk = ctx.sample_packets(n, seed=42)
        d = pk["dir"]
        # test/testPhotonSource.cpp:92-131: isotropy
        assert np.abs(d.mean(axis=0)).max() < 3e-3
        assert np.abs(np.linalg.norm(d, axis=1) - 1.).max() < 1e-14
        assert abs((d[:, 2] ** 2).mean() - 1. / 3.) < 2e-3
        # tau ~ Exp(1)
        assert abs(pk["tau"].mean() - 1.) < 5e-3 and abs(pk["tau"].var() - 1.) < 2e-2
        # Planck sampler vs the reference's sampler (its own RNG): compare histograms
        nu_ref = ref.sample_spectrum(0, 40000., n, seed=42)
        edges = np.linspace(3.288465385e15, 4 * 3.288465385e15, 101)
        h1, _ = np.histogram(pk["nu"], edges)
        h2, _ = np.histogram(nu_ref, edges)
        big = h2 > 2000
        assert np.abs(h1[big] - h2[big]).max() / np.sqrt(2 * h2[big]).max() < 6.
        assert np.abs((h1[big] - h2[big]) / np.sqrt(h1[big] + h2[big])).max() < 5.
        # streams do not depend on how a batch is split (multi-GPU sharding)
        a = ctx.sample_packets(1000, offset=0, seed=9)
        b = ctx.sample_packets(500, offset=500, seed=9)
        assert np.array_equal(a["dir"][500:], b["dir"]) and np.array_equal(a["nu"][500:], b["nu"])
        c = ctx.sample_packets(1000, offset=0, seed=10)
        assert not np.array_equal(a["dir"], c["dir"])


def test_diffuse_spectra_statistics(cmib, ref):
    from cmacionize_b200 import capi
    with cmib.Context([-1, -1, -1], [2, 2, 2], [4, 4, 4]) as ctx:
        ctx.set_abundances(*ABUNDANCES)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_reemission(capi.REEMISSION_PHYSICAL)
        # tables built on the host must be the reference's tables bit for bit
        f, t, c = ctx.get_spectrum_tables(1)
        rf, rt, rc = ref.lyc_tables(0, 1)
        assert np.array_equal(f, rf) and np.array_equal(t, rt) and np.array_equal(c, rc)
        f, t, c = ctx.get_spectrum_tables(2)
        rf, rt, rc = ref.lyc_tables(1, 1)
        assert np.array_equal(c, rc)
        f, c = ctx.get_spectrum_tables(3)
        rf, rc = ref.he2pc_tables()
        assert np.array_equal(f, rf) and np.array_equal(c, rc)
        n = 400_000
        for which, T in ((1, 8000.), (2, 8000.), (3, 0.), (1, 1000.), (2, 20000.)):
            a = ctx.sample_spectrum(which, T, n, seed=3)
            b = ref.sample_spectrum(which, T, n, seed=3)
            lo, hi = min(a.min(), b.min()), max(a.max(), b.max())
            h1, e = np.histogram(a, 60, (lo, hi))
            h2, _ = np.histogram(b, 60, (lo, hi))
            big = (h1 + h2) > 400
            z = (h1[big] - h2[big]) / np.sqrt(h1[big] + h2[big])
            assert np.abs(z).max() < 5., (which, T, np.abs(z).max())


def test_shoot_is_deterministic_and_shardable(cmib):
    """Same seed -> same accumulators (up to atomic summation order); [0,N) in one call ==
    [0,N/2) + [N/2,N) in two calls: the property the 1/2/4/8-GPU split relies on."""
    from cmacionize_b200 import problems
    prob = problems.stromgren(ncell=32, n_packets=100000)
    ctx = prob.ctx
    outs = []
    for split in (1, 1, 4):
        ctx.reset_accumulators()
        per = 100000 // split
        tot = 0.
        for k in range(split):
            tw, tc = ctx.shoot(per, packet_offset=k * per, seed=5, iteration=2)
            tot += tw
        assert tot == 100000.
        outs.append(ctx.download_accumulators()[0][0].copy())
    s = outs[0].max()
    assert np.abs(outs[0] - outs[1]).max() <= 1e-12 * s
    assert np.abs(outs[0] - outs[2]).max() <= 1e-12 * s
    ctx.reset_accumulators()
    ctx.shoot(100000, seed=6, iteration=2)
    other = ctx.download_accumulators()[0][0]
    assert np.abs(outs[0] - other).max() > 1e-6 * s  # a different seed is a different sample
    ctx.close()


@pytest.mark.parametrize("config", ["stromgren", "stromgren_diffuse", "lexington", "fixed_reemission", "periodic",
                                    "continuous", "continuous_only", "planar", "distant_star", "extended_disc", "spiral_galaxy", "bimodal",
                                    "periodic_honly", "zero_direction"])
def test_wavefront_pipeline_equals_the_per_packet_kernel(cmib, config):
    """The production shoot (prepare/march kernels + device queues, wavefront.cuh) and the
    one-thread-per-packet kernel run the same shoot_packet logic on the same per-packet random
    streams: identical counters (packets by type, cell crossings, (re-)emissions) and
    accumulators equal up to the order of the atomic adds.  Small queue capacities force many
    rounds, chunk boundaries and partially filled warps; the queue orders (1: ordered queue read by the
    plain kernel, 2: the coherent march = counting sort by (source | direction | depth) + in-warp sums of
    same-cell terms), two lanes and the round-by-round tail only change which packets run together and in which order their terms are added."""
    import os
    from cmacionize_b200 import problems, capi
    npk = 60000
    if config == "stromgren":
        prob = problems.stromgren(ncell=32, n_packets=npk)
    elif config == "stromgren_diffuse":
        prob = problems.stromgren(ncell=32, n_packets=npk, diffuse=True)
    elif config == "lexington":
        prob = problems.lexington(20, ncell=24, n_packets=npk)
    elif config == "fixed_reemission":
        prob = problems.stromgren(ncell=16, n_packets=npk)
        prob.ctx.set_reemission(capi.REEMISSION_FIXED_VALUE, 0.364, problems.ev_to_hz(19.8))
    elif config == "continuous":
        # star + isotropic external field (half of the packets each, the external ones weigh
        # L_c / L_d = 0.25 and keep that weight through their re-emissions)
        prob = problems.stromgren(ncell=32, n_packets=npk, diffuse=True)
        prob.ctx.set_continuous_source(capi.CONTINUOUS_ISOTROPIC, 0.25 * 4.26e49, capi.SPECTRUM_MONOCHROMATIC,
                                       problems.ev_to_hz(13.6))
    elif config == "planar":
        # PlanarContinuousPhotonSource: an emitting sheet on a cell boundary (packets start ON a wall)
        prob = problems.stromgren(ncell=32, n_packets=npk, diffuse=True)
        prob.ctx.set_sources(None, None, 0.)
        prob.ctx.set_planar_source_geometry(2, 0., [-4 * PC, -5 * PC], [8 * PC, 10 * PC])
        prob.ctx.set_continuous_source(capi.CONTINUOUS_PLANAR, 3e49, capi.SPECTRUM_MONOCHROMATIC, problems.ev_to_hz(13.6))
    elif config == "distant_star":
        # DistantStarContinuousPhotonSource: rejection sampling, a variable number of deviates per packet
        prob = problems.stromgren(ncell=32, n_packets=npk, diffuse=True)
        prob.ctx.set_distant_star_position([-9 * PC, -7 * PC, 1 * PC])
        prob.ctx.set_continuous_source(capi.CONTINUOUS_DISTANT_STAR, 0.5 * 4.26e49, capi.SPECTRUM_MONOCHROMATIC,
                                       problems.ev_to_hz(13.6))
    elif config == "extended_disc":
        # ExtendedDiscContinuousPhotonSource: emission from a volume, the Gaussian height redrawn while outside the box
        prob = problems.stromgren(ncell=32, n_packets=npk, diffuse=True)
        prob.ctx.set_sources(None, None, 0.)
        prob.ctx.set_extended_disc_geometry(1, 3 * PC, 1.5 * PC)
        prob.ctx.set_continuous_source(capi.CONTINUOUS_EXTENDED_DISC, 3e49, capi.SPECTRUM_MONOCHROMATIC, problems.ev_to_hz(13.6))
    elif config == "spiral_galaxy":
        # SpiralGalaxyContinuousPhotonSource: bulge / disc positions redrawn until inside the box (the bulge radii are
        # hard-wired kpc, so in this pc-sized box every bulge draw is rejected: a variable number of deviates per packet)
        prob = problems.stromgren(ncell=32, n_packets=npk, diffuse=True)
        prob.ctx.set_sources(None, None, 0.)
        prob.ctx.set_spiral_galaxy_geometry(3 * PC, 0.5 * PC, 0.001)
        prob.ctx.set_continuous_source(capi.CONTINUOUS_SPIRAL_GALAXY, 3e49, capi.SPECTRUM_MONOCHROMATIC, problems.ev_to_hz(13.6))
    elif config == "bimodal":
        # CrossSections: Bimodal (two constant values per ion, split at a frequency limit) with the diffuse field
        prob = problems.lexington(20, ncell=24, n_packets=npk)
        low, high = np.zeros(14), np.zeros(14)
        low[0], high[0] = 6.3e-22, 2.5e-22
        high[1] = 7.e-22
        low[7], high[7], low[4], high[9] = 3e-22, 1e-22, 2e-22, 4e-22
        prob.ctx.set_bimodal_cross_sections(problems.ev_to_hz(20.), low, high)
        prob.ctx.set_reemission(capi.REEMISSION_PHYSICAL)        # rebuilds its tables from these cross sections
        nu_probe = problems.ev_to_hz(np.array([13.7, 19.99, 20.0, 30.]))
        sig = prob.ctx.eval_cross_sections(nu_probe)
        assert np.array_equal(sig[:2], np.tile(low, (2, 1))) and np.array_equal(sig[2:], np.tile(high, (2, 1)))
    elif config in ("periodic_honly", "zero_direction"):
        # H-only layout (FixedValue cross sections) in a non-cubic box, two sources, re-emission at a fixed frequency
        # (so there is a heat term): with two periodic axes ("periodic_honly") and without ("zero_direction" is
        # that twin; explicit zero direction components are walked in test_gpu_march.py) — the periodic + heat and
        # the heat variants of the coherent H-only walk (march_lean_kernel<true, *, ..>)
        per = (True, False, True) if config == "periodic_honly" else (False, False, False)
        ctx = cmib.Context([-1e17, -2e17, -3e17], [2e17, 5e17, 3e17], [12, 20, 9], periodic=per)
        ctx.set_abundances()
        sig = np.zeros(capi.NUM_IONS); sig[0] = 6.3e-22
        ctx.set_cross_sections(capi.CROSS_SECTIONS_FIXED_VALUE, sig)
        rr = np.zeros(capi.NUM_IONS); rr[0] = 4.e-19
        ctx.set_recombination_rates(capi.RECOMBINATION_FIXED_VALUE, rr)
        ctx.set_sources([[0., 0., 0.], [0.9e17, 2.9e17, -2.9e17]], [0.6, 0.4], 1e49)
        ctx.set_spectrum(capi.SPECTRUM_MONOCHROMATIC, problems.ev_to_hz(15.))
        ctx.set_reemission(capi.REEMISSION_FIXED_VALUE, 0.4, problems.ev_to_hz(14.2))
        ctx.set_temperature_params(do_temperature_calculation=False)
        nc = ctx.ncells
        prob = problems.Problem(config, ctx, np.full(nc, 3e8), np.full(nc, 8000.), problems.initial_fractions(nc), npk, 1)
        prob.upload()
    elif config == "continuous_only":
        prob = problems.lexington(20, ncell=24, n_packets=npk)
        prob.ctx.set_sources(None, None, 0.)
        # a tabulated spectrum (the form of FaucherGiguere, WMBasic, ...: cmib_set_spectrum_table)
        trng = np.random.default_rng(8)
        cdf = np.concatenate([[0.], np.cumsum(trng.uniform(0.1, 1., 99) * np.exp(-np.arange(99) / 30.))])
        prob.ctx.set_spectrum_table(np.linspace(3.289e15, 4 * 3.289e15, 100), cdf / cdf[-1], role=1)
        prob.ctx.set_continuous_source(capi.CONTINUOUS_ISOTROPIC, 1e49, capi.SPECTRUM_TABULATED)
    else:
        # non-cubic box, periodic in x and z, three sources (one outside the box: its packets
        # are lost immediately, as in the reference), Physical re-emission
        ctx = cmib.Context([-1e17, -2e17, -3e17], [2e17, 5e17, 3e17], [12, 20, 9], periodic=(True, False, True))
        ctx.set_abundances(He=0.1)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_recombination_rates(capi.RECOMBINATION_VERNER)
        ctx.set_sources([[0., 0., 0.], [0.9e17, 2.9e17, -2.9e17], [0., 9e17, 0.]], [0.5, 0.4, 0.1], 1e49)
        ctx.set_spectrum(capi.SPECTRUM_PLANCK, 30000.)
        ctx.set_reemission(capi.REEMISSION_PHYSICAL)
        nc = ctx.ncells
        prob = problems.Problem("periodic", ctx, np.full(nc, 3e8), np.full(nc, 8000.), problems.initial_fractions(nc),
                                npk, 1)
        prob.upload()
    ctx = prob.ctx
    rng = np.random.default_rng(17)
    x = prob.ionic_fractions.copy()
    x[0] = np.exp(rng.uniform(np.log(1e-5), np.log(1e-2), ctx.ncells))
    x[1] = np.exp(rng.uniform(np.log(1e-5), np.log(1e-1), ctx.ncells))
    ctx.upload_cells(prob.number_density, np.where(prob.number_density > 0, 7500., 0.), x)
    results = []
    # sort 1 = ordered queue read by the plain kernel, 2 = the coherent march (ordered queue + in-warp sums);
    # lanes 2 = two sets of queues on two streams, out of phase; tail 0 = the last generations of re-emitted packets
    # round by round instead of in the tail kernel
    for algorithm, capacity, sort, lanes, tail in ((1, None, 0, 1, 1), (0, None, 0, 1, 1), (0, 4096, 0, 1, 1), (0, 1024, 0, 1, 0),
                                                   (0, None, 1, 1, 1), (0, 2048, 1, 1, 1), (0, None, 2, 1, 1), (0, 2048, 2, 1, 1),
                                                   (0, 1024, 2, 1, 0), (0, 2048, 0, 2, 1), (0, 2048, 2, 2, 1), (0, None, 2, 2, 0)):
        if capacity is None:
            os.environ.pop("CMIB_QUEUE_CAPACITY", None)
        else:
            os.environ["CMIB_QUEUE_CAPACITY"] = str(capacity)
        os.environ["CMIB_SORT"] = str(sort)   # coherence sort of the march queue (wavefront.cuh)
        os.environ["CMIB_LANES"] = str(lanes)
        os.environ["CMIB_TAIL"] = str(tail)
        ctx.set_shoot_algorithm(algorithm)
        ctx.reset_accumulators()
        ctx.update_reemission_probabilities()
        tw, tc = ctx.shoot(npk, packet_offset=1000, seed=99, iteration=4)
        J, heat = ctx.download_accumulators()
        results.append((tw, tc, ctx.shoot_statistics(), J, heat))
    for k in ("CMIB_QUEUE_CAPACITY", "CMIB_SORT", "CMIB_LANES", "CMIB_TAIL"):
        os.environ.pop(k, None)
    ctx.close()
    tw0, tc0, st0, J0, h0 = results[0]
    if config in ("continuous", "distant_star"):
        wc = 0.25 if config == "continuous" else 0.5
        ncont = round((npk - tw0) / (1. - wc))  # tw = n_discrete + wc n_continuous
        assert abs(ncont - 0.5 * npk) < 5 * np.sqrt(0.25 * npk) and abs(tc0.sum() - tw0) < 1e-9 * tw0
    else:
        assert tw0 == npk and tc0.sum() == npk
    if config == "periodic":
        assert 0.05 * npk < tc0[0] < 0.5 * npk   # the 10 % emitted outside + escapes through the y faces
    if config != "stromgren":
        assert st0[1] > 1.05 * npk  # re-emission happened
    for tw, tc, st, J, h in results[1:]:
        assert abs(tw - tw0) <= 1e-12 * tw0 and np.abs(tc - tc0).max() <= 1e-12 * tw0 and st == st0
        for k in range(14):
            assert np.abs(J[k] - J0[k]).max() <= 1e-12 * max(J0[k].max(), 1e-300), k
        for k in range(2):
            assert np.abs(h[k] - h0[k]).max() <= 1e-12 * max(np.abs(h0[k]).max(), 1e-300), k


@pytest.mark.parametrize("config", ["stromgren64_1e6", "lexingtonHII20_64_1e7"])
def test_full_size_checksum_of_the_accumulation(cmib, config):
    """BASELINE-size runs cannot be replayed by the CPU oracle in seconds; they are pinned through an
    identity that holds for any grid and any number of packets: every crossing adds
    ds w sigma_k to J_k of its cell and removes ds n (sigma_H x_H + A_He sigma_He x_He) from the
    packet's optical depth, so with unit weights

        sum_cells n (x_H J_H + A_He x_He J_He)  ==  optical depth traversed by all packets,

    the right side being summed per packet inside the walk (cmib_shoot_optical_depth).  A lost or
    mis-addressed atomic add breaks it (cells differ in n and x).  Plus packet conservation."""
    from cmacionize_b200 import problems
    if config == "stromgren64_1e6":
        prob = problems.stromgren(ncell=64, n_packets=1_000_000)
        A_He, npk, spin = 0., 1_000_000, 3
    else:
        prob = problems.lexington(20, ncell=64, n_packets=10_000_000)
        A_He, npk, spin = 0.1, 10_000_000, 5
    ctx = prob.ctx
    for loop in range(spin):
        problems.run_iteration(prob, loop, n_packets=1_000_000)
    # make the cells pairwise different so that a mis-addressed add cannot cancel
    n, T, x, _ = ctx.download_cells()
    rng = np.random.default_rng(1)
    n = n * rng.uniform(0.5, 1.5, n.size)
    ctx.upload_cells(n, T, x)
    ctx.reset_accumulators()
    ctx.update_reemission_probabilities()
    tw, tc = ctx.shoot(npk, seed=11, iteration=spin)
    crossings, emissions = ctx.shoot_statistics()
    tau = ctx.shoot_optical_depth()
    J, heat = ctx.download_accumulators()
    ctx.close()
    assert tw == npk and tc.sum() == npk
    assert emissions >= npk and crossings > 10 * npk
    lhs = float(np.sum(n * (x[0] * J[0] + A_He * x[1] * J[1])))
    assert abs(lhs - tau) <= 1e-9 * tau, (lhs, tau)
    # mean optical depth per emission is ~1 for absorbed packets and < 1 for escaping ones
    assert 0.05 < tau / emissions < 1.05
    assert (J[0][n > 0] > 0).mean() > 0.2 and (heat[0] >= 0).all()


def test_continuous_source_shoot_equals_oracle(cmib, ref):
    """Star + IsotropicContinuousPhotonSource (PhotonSource.cpp:113-131, 208-249: half of the packets
    each, an external packet weighs L_c / L_d).  The exported packets, with their weights, pushed
    through the oracle's interact() give shoot's accumulators; the external packets start on the
    faces of the box, point inwards, and are distributed like the reference's own sampler's."""
    from cmacionize_b200 import capi
    anchor, sides, nc3 = [-3 * PC, -2 * PC, -1 * PC], [6 * PC, 4 * PC, 2 * PC], [24, 16, 8]
    with cmib.Context(anchor, sides, nc3) as ctx:
        nc = ctx.ncells
        ctx.set_abundances(*ABUNDANCES)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_sources([[0., 0., 0.]], [1.], 1e49)
        ctx.set_spectrum(capi.SPECTRUM_PLANCK, 40000.)
        ctx.set_continuous_source(capi.CONTINUOUS_ISOTROPIC, 3e48, capi.SPECTRUM_PLANCK, 25000.)
        rng = np.random.default_rng(5)
        n = 1e8 * np.exp(rng.normal(0, 0.5, nc))
        x = np.zeros((14, nc))
        x[0] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), nc))
        x[1] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), nc))
        ctx.upload_cells(n, np.full(nc, 8000.), x)
        ctx.reset_accumulators()
        npk = 40000
        tw, tc = ctx.shoot(npk, seed=11, iteration=0)
        J, heat = ctx.download_accumulators()
        pk = ctx.sample_packets(npk, seed=11, iteration=0)
        external = ~np.all(pk["pos"] == 0., axis=1)
        w = np.where(external, 0.3, 1.)
        assert abs(external.mean() - 0.5) < 5 * 0.5 / np.sqrt(npk)
        assert abs(tw - w.sum()) <= 1e-9 * tw and abs(tc.sum() - tw) <= 1e-9 * tw
        r = ref.interact(anchor, sides, nc3, [0, 0, 0], n, x[0], x[1], pk["pos"], pk["dir"], pk["sigma"],
                         pk["sigma_He_corr"], pk["nu"], w, pk["tau"])
        for k in range(14):
            assert np.abs(J[k] - r["J"][k]).max() <= 1e-11 * max(r["J"][k].max(), 1e-300), k
        for k in range(2):
            assert np.abs(heat[k] - r["heat"][k]).max() <= 1e-11 * np.abs(r["heat"][k]).max()
        # the two spectra: external packets are softer (25000 K vs 40000 K)
        assert pk["nu"][external].mean() < 0.9 * pk["nu"][~external].mean()
        # external packets only, against the reference's sampler
        ctx.set_sources(None, None, 0.)
        m = 400000
        pk = ctx.sample_packets(m, seed=3)
        a, sd = np.array(anchor), np.array(sides)
        u, rpos, rdir = ref.isotropic_incoming(anchor, sides, m, seed=3)

        def face_stats(pos, d):
            lo = np.isclose(pos, a, rtol=0, atol=1e-9 * sd)
            hi = np.isclose(pos, a + sd, rtol=0, atol=1e-9 * sd)
            assert (lo | hi).any(axis=1).all()
            inward = np.where(lo, d, np.where(hi, -d, np.nan))
            assert np.nanmin(inward) >= 0.
            frac = np.array([(lo[:, k] | hi[:, k]).mean() for k in range(3)])
            mu = np.array([np.nanmean(inward[:, k]) for k in range(3)])
            return frac, mu

        assert (pk["pos"] >= a).all() and (pk["pos"] < a + sd).all()
        f1, mu1 = face_stats(pk["pos"], pk["dir"])
        f2, mu2 = face_stats(rpos, rdir)
        assert np.abs(f1 - f2).max() < 5 * np.sqrt(2 * 0.25 / m)
        assert np.abs(mu1 - mu2).max() < 5e-3
        assert np.abs(np.linalg.norm(pk["dir"], axis=1) - 1.).max() < 1e-14


def test_tabulated_and_uniform_spectra_on_the_device(cmib, ref):
    """cmib_set_spectrum_table with the reference's FaucherGiguere tables (z = 7) and the Uniform
    spectrum, sampled on the device, against the reference's own samplers (histograms; the sampling
    rule itself is pinned bit for bit on the CPU tier, test_host_physics.py)."""
    from cmacionize_b200 import capi
    n = 1_000_000
    with cmib.Context([-1, -1, -1], [2, 2, 2], [4, 4, 4]) as ctx:
        ctx.set_abundances(*ABUNDANCES)
        ctx.set_cross_sections(capi.CROSS_SECTIONS_VERNER)
        ctx.set_sources([[0., 0., 0.]], [1.], 1e49)
        d = ref.faucher_giguere(7., n, seed=42)
        for role, which in ((0, 0), (1, 4)):
            ctx.set_spectrum_table(d["freq"], d["cdf"], role=role)
            nu = ctx.sample_spectrum(which, 0., n, seed=3)
            assert nu.min() >= d["freq"][0] and nu.max() <= d["freq"][-1]
            edges = np.linspace(d["freq"][0], d["freq"][-1], 81)
            h1, _ = np.histogram(nu, edges)
            h2, _ = np.histogram(d["nu"], edges)
            big = h2 > 1000
            assert big.sum() > 20
            assert np.abs((h1[big] - h2[big]) / np.sqrt(h1[big] + h2[big])).max() < 5.
        ctx.set_spectrum(capi.SPECTRUM_UNIFORM, 0.)
        nu = ctx.sample_spectrum(0, 0., n, seed=4)
        _, nu_ref = ref.uniform_spectrum(n, seed=4)
        assert nu.min() >= 3.289e15 and nu.max() <= 4 * 3.289e15
        assert abs(nu.mean() / nu_ref.mean() - 1.) < 2e-3 and abs(nu.std() / nu_ref.std() - 1.) < 3e-3
        pk = ctx.sample_packets(1000, seed=1)   # emission uses it
        assert (pk["nu"] >= 3.289e15).all() and (pk["nu"] <= 4 * 3.289e15).all() and np.unique(pk["nu"]).size > 990


@pytest.mark.parametrize("capacity", [None, 262144])
def test_measured_queue_order_changes_nothing(cmib, capacity):
    """Grids that do not fit in L2 choose the order of the march queue by measurement (cmib_api.cu,
    sort_mode -1): a warm-up shoot, then the two orders timed either on two successive shoots or
    — when a shoot holds at least four queue capacities — on rounds 1 and 2 of one shoot, then the
    winner.  Whatever it picks, every shoot must give the sums of the plain order on the same
    packets (order of the atomic adds aside)."""
    import os
    from cmacionize_b200 import problems
    npk = 1_200_000
    prob = problems.lexington(20, ncell=128, n_packets=npk)      # 128^3 x 160 B = 335 MB > L2
    ctx = prob.ctx
    rng = np.random.default_rng(3)
    x = prob.ionic_fractions.copy()
    x[0] = np.exp(rng.uniform(np.log(1e-5), np.log(1e-3), ctx.ncells))
    x[1] = np.exp(rng.uniform(np.log(1e-5), np.log(1e-2), ctx.ncells))
    ctx.upload_cells(prob.number_density, np.where(prob.number_density > 0, 7500., 0.), x)
    ctx.update_reemission_probabilities()
    if capacity is None:
        os.environ.pop("CMIB_QUEUE_CAPACITY", None)
    else:
        os.environ["CMIB_QUEUE_CAPACITY"] = str(capacity)
    try:
        for it in range(5):
            out = []
            for order in (None, "0"):
                if order is None:
                    os.environ.pop("CMIB_SORT", None)
                else:
                    os.environ["CMIB_SORT"] = order
                ctx.reset_accumulators()
                tw, tc = ctx.shoot(npk, seed=5, iteration=it)
                J, heat = ctx.download_accumulators()
                out.append((tw, tc, ctx.shoot_statistics(), J, heat))
            (tw, tc, st, J, h), (tw0, tc0, st0, J0, h0) = out
            assert tw == tw0 and np.array_equal(tc, tc0) and st == st0, it
            for k in range(14):
                assert np.abs(J[k] - J0[k]).max() <= 1e-12 * max(J0[k].max(), 1e-300), (it, k)
            for k in range(2):
                assert np.abs(h[k] - h0[k]).max() <= 1e-12 * max(np.abs(h0[k]).max(), 1e-300), (it, k)
    finally:
        os.environ.pop("CMIB_QUEUE_CAPACITY", None)
        os.environ.pop("CMIB_SORT", None)
        ctx.close()


def test_source_configuration_errors(cmib):
    """Misuse is reported with the reference's messages (and never computes): no sources at all, a
    distant star inside the box (DistantStarContinuousPhotonSource.hpp:78-81), geometry forgotten,
    unknown types, a non-monotonic tabulated spectrum, zero luminosity."""
    from cmacionize_b200 import capi
    with cmib.Context([-1., -1., -1.], [2., 2., 2.], [4, 4, 4]) as ctx:
        with pytest.raises(cmib.CmibError, match="no photon sources set"):
            ctx.shoot(10)
        with pytest.raises(cmib.CmibError, match="lies inside the simulation box"):
            ctx.set_distant_star_position([0.5, 0., 0.])
        with pytest.raises(cmib.CmibError, match="cmib_set_distant_star_position before"):
            ctx.set_continuous_source(capi.CONTINUOUS_DISTANT_STAR, 1e49, capi.SPECTRUM_MONOCHROMATIC, 3.3e15)
        with pytest.raises(cmib.CmibError, match="cmib_set_planar_source_geometry before"):
            ctx.set_continuous_source(capi.CONTINUOUS_PLANAR, 1e49, capi.SPECTRUM_MONOCHROMATIC, 3.3e15)
        with pytest.raises(cmib.CmibError, match="cmib_set_extended_disc_geometry before"):
            ctx.set_continuous_source(capi.CONTINUOUS_EXTENDED_DISC, 1e49, capi.SPECTRUM_MONOCHROMATIC, 3.3e15)
        with pytest.raises(cmib.CmibError, match="more than 10 scale heights outside"):
            ctx.set_extended_disc_geometry(2, 5., 0.1)
        with pytest.raises(cmib.CmibError, match="scale height of the disc must be positive"):
            ctx.set_extended_disc_geometry(2, 0., 0.)
        with pytest.raises(cmib.CmibError, match="cmib_set_spiral_galaxy_geometry before"):
            ctx.set_continuous_source(capi.CONTINUOUS_SPIRAL_GALAXY, 1e49, capi.SPECTRUM_MONOCHROMATIC, 3.3e15)
        with pytest.raises(cmib.CmibError, match="bulge over total ratio must lie in"):
            ctx.set_spiral_galaxy_geometry(1., 0.1, 1.5)
        with pytest.raises(cmib.CmibError, match="Unknown ContinuousPhotonSource type"):
            ctx.set_continuous_source(7, 1e49, capi.SPECTRUM_MONOCHROMATIC, 3.3e15)
        with pytest.raises(cmib.CmibError, match="Unknown PhotonSourceSpectrum type"):
            ctx.set_spectrum(9, 0.)
        with pytest.raises(cmib.CmibError, match="positive luminosity"):
            ctx.set_continuous_source(capi.CONTINUOUS_ISOTROPIC, 0., capi.SPECTRUM_MONOCHROMATIC, 3.3e15)
        with pytest.raises(cmib.CmibError, match="must not decrease"):
            ctx.set_spectrum_table([3.3e15, 4e15, 5e15], [0., 0.7, 0.5])
        with pytest.raises(cmib.CmibError, match="at least two frequencies"):
            ctx.set_spectrum_table([3.3e15], [1.])
        with pytest.raises(cmib.CmibError, match="do not sum to 1"):
            ctx.set_sources([[0., 0., 0.], [0.1, 0., 0.]], [0.5, 0.4], 1e49)


def test_task_based_packet_conventions(cmib):
    """cmib_set_packet_conventions(TASK_BASED): the packets carry abundance-weighted cross sections and the heating
    terms use the hard-coded thresholds of DensitySubGrid.hpp:608-612.  Same packets as under the default convention
    (same Philox streams), so per cell  J_ion(task) = A_element(ion) * J_ion(default),  heat_H(task) = heat_H(default) +
    (nu_H(13.6 eV) - 3.288e15) * J_H,  heat_He(task) = A_He * (heat_He(default) + (nu_He(24.6 eV) - 5.948e15) * J_He);
    and the state update divides the abundances out again (TaskBasedIonizationSimulation.cpp:932-951): the
    ionization-only update gives the same fractions."""
    from cmacionize_b200 import problems
    npk = 400_000
    out = []
    for conv in (0, 1):
        prob = problems.lexington(20, ncell=16, n_packets=npk)
        ctx = prob.ctx
        ctx.set_packet_conventions(conv)
        rng = np.random.default_rng(5)
        x = prob.ionic_fractions.copy()
        x[0] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-2), ctx.ncells))
        x[1] = np.exp(rng.uniform(np.log(1e-4), np.log(1e-1), ctx.ncells))
        ctx.upload_cells(prob.number_density, np.where(prob.number_density > 0, 7500., 0.), x)
        ctx.reset_accumulators()
        ctx.update_reemission_probabilities()
        tw, tc = ctx.shoot(npk, seed=77, iteration=1)
        J, heat = ctx.download_accumulators()
        ctx.update_state(1, 0.)
        n1, T1, x1, h1 = ctx.download_cells()
        out.append((tw, tc.copy(), J, heat, x1, h1))
        ctx.close()
    (tw0, tc0, J0, H0, x0, h0), (tw1, tc1, J1, H1, x1, h1) = out
    assert tw0 == tw1 and np.array_equal(tc0, tc1)            # every packet met the same fate
    A = dict(He=0.1, C=2.2e-4, N=4.e-5, O=3.3e-4, Ne=5.e-5, S=9.e-6)
    el = ["H", "He", "C", "C", "N", "N", "N", "O", "O", "Ne", "Ne", "S", "S", "S"]
    assert np.abs(J1[0] - J0[0]).max() <= 1e-12 * J0[0].max()
    for ion in range(1, 14):
        a = A[el[ion]]
        assert np.abs(J1[ion] - a * J0[ion]).max() <= 1e-12 * a * J0[ion].max(), ion
    assert sum(J0[ion].max() > 0. for ion in range(14)) >= 8     # the hard ions see no photon of a 20 000 K star
    nuH, nuHe = problems.ev_to_hz(13.6), problems.ev_to_hz(24.6)
    eH = H0[0] + (nuH - 3.288e15) * J0[0]
    eHe = A["He"] * (H0[1] + (nuHe - 5.948e15) * J0[1])
    assert np.abs(H1[0] - eH).max() <= 1e-10 * np.abs(eH).max()
    assert np.abs(H1[1] - eHe).max() <= 1e-10 * np.abs(eHe).max()
    ok = np.isfinite(x0)
    assert np.array_equal(ok, np.isfinite(x1))
    assert np.abs(x1[ok] - x0[ok]).max() <= 1e-10               # ionization-only update: J / A * A to rounding


def test_shoot_of_more_rounds_than_one_group_of_rounds(cmib):
    """A shoot whose primaries need more rounds than the host enqueues in one group (63): 300 rounds of 1024 packets
    without re-emission, and the same with re-emission (groups of 4 kept ahead of the read-back), against one round."""
    import os
    from cmacionize_b200 import problems
    npk = 300_000
    for diffuse in (False, True):
        prob = problems.stromgren(ncell=16, n_packets=npk, diffuse=diffuse)
        ctx = prob.ctx
        out = []
        for cap in (None, 1024):
            if cap is None:
                os.environ.pop("CMIB_QUEUE_CAPACITY", None)
            else:
                os.environ["CMIB_QUEUE_CAPACITY"] = str(cap)
            try:
                ctx.reset_accumulators()
                ctx.update_reemission_probabilities()
                tw, tc = ctx.shoot(npk, seed=3, iteration=1)
            finally:
                os.environ.pop("CMIB_QUEUE_CAPACITY", None)
            J, heat = ctx.download_accumulators()
            out.append((tw, tc.copy(), ctx.shoot_statistics(), J))
        ctx.close()
        (tw0, tc0, st0, J0), (tw1, tc1, st1, J1) = out
        assert tw0 == tw1 == npk and np.array_equal(tc0, tc1) and st0 == st1
        assert np.abs(J1[0] - J0[0]).max() <= 1e-12 * J0[0].max()
