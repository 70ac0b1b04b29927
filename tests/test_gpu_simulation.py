"""GPU tier, tier 3 of the north star: converged fields against the reference's own
CPU run of the same problem, within Monte Carlo noise.

The noise tolerance is defined the way SURVEY.md §7 (hard part 4) prescribes: from
two reference runs with different seeds.  The GPU run must be no further from a
reference run than two reference runs are from each other (times a small factor).
The two reference seeds are far apart on purpose: the reference seeds OpenMP thread t with
seed + t (IonizationPhotonShootJobMarket.hpp:80-87), so runs with seeds 42 and 43 share all
but one of their random streams and under-estimate the noise."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

PC = 3.086e16

STROMGREN_PARAM = """
SimulationBox:
  anchor: [-5. pc, -5. pc, -5. pc]
  sides: [10. pc, 10. pc, 10. pc]
  periodicity: [false, false, false]
DensityGrid:
  type: Cartesian
  number of cells: [{nc}, {nc}, {nc}]
DensityFunction:
  type: Homogeneous
  density: 100. cm^-3
  temperature: 8000. K
TemperatureCalculator:
  do temperature calculation: false
PhotonSourceDistribution:
  type: SingleStar
  position: [0. pc, 0. pc, 0. pc]
  luminosity: 4.26e49 s^-1
PhotonSourceSpectrum:
  type: Monochromatic
  frequency: 13.6 eV
IonizationSimulation:
  number of photons: {npk}
  number of iterations: {nit}
  random seed: {seed}
CrossSections:
  type: FixedValue
  hydrogen_0: 6.3e-18 cm^2
  helium_0: 0. m^2
  carbon_1: 0. m^2
  carbon_2: 0. m^2
  nitrogen_0: 0. m^2
  nitrogen_1: 0. m^2
  nitrogen_2: 0. m^2
  oxygen_0: 0. m^2
  oxygen_1: 0. m^2
  neon_0: 0. m^2
  neon_1: 0. m^2
  sulphur_1: 0. m^2
  sulphur_2: 0. m^2
  sulphur_3: 0. m^2
RecombinationRates:
  type: FixedValue
  hydrogen_1: 4.e-13 cm^3 s^-1
  helium_1: 0. m^3 s^-1
  carbon_2: 0. m^3 s^-1
  carbon_3: 0. m^3 s^-1
  nitrogen_1: 0. m^3 s^-1
  nitrogen_2: 0. m^3 s^-1
  nitrogen_3: 0. m^3 s^-1
  oxygen_1: 0. m^3 s^-1
  oxygen_2: 0. m^3 s^-1
  neon_1: 0. m^3 s^-1
  neon_2: 0. m^3 s^-1
  sulphur_2: 0. m^3 s^-1
  sulphur_3: 0. m^3 s^-1
  sulphur_4: 0. m^3 s^-1
{extra}
"""


def radial_profile(x, nc, half):
    cs = 2 * half / nc
    m = -half + cs * (np.arange(nc) + 0.5)
    X, Y, Z = np.meshgrid(m, m, m, indexing="ij")
    r = np.sqrt(X * X + Y * Y + Z * Z).reshape(-1)
    return r


def stromgren_radius(xH, r):
    """radius where the spherically averaged neutral fraction crosses 0.5"""
    bins = np.linspace(0, r.max(), 80)
    idx = np.digitize(r, bins)
    prof = np.array([xH[idx == i].mean() if (idx == i).any() else np.nan for i in range(1, len(bins))])
    mid = 0.5 * (bins[1:] + bins[:-1])
    k = np.where(prof > 0.5)[0][0]
    return np.interp(0.5, [prof[k - 1], prof[k]], [mid[k - 1], mid[k]])


def shell_means(f, r, edges):
    idx = np.digitize(r, edges)
    return np.array([f[idx == i].mean() for i in range(1, len(edges))])


@pytest.mark.parametrize("diffuse", [False, True])
def test_stromgren_converges_to_the_reference(cmib, ref, tmp_path, diffuse):
    from cmacionize_b200 import problems
    nc, npk, nit = 32, 1000000, 10
    extra = "DiffuseReemissionHandler:\n  type: Physical\n" if diffuse else ""
    runs = []
    for seed in (42, 4242):
        pf = tmp_path / f"stromgren_{seed}.param"
        pf.write_text(STROMGREN_PARAM.format(nc=nc, npk=npk, nit=nit, seed=seed, extra=extra))
        fields, _ = ref.run_paramfile(pf, nc ** 3)
        runs.append(fields)
    prob = problems.stromgren(ncell=nc, n_packets=npk, n_iterations=nit, diffuse=diffuse)
    problems.run(prob)
    n, T, x, heat = prob.ctx.download_cells()
    prob.ctx.close()
    assert np.array_equal(n, runs[0][0]) and np.array_equal(T, runs[0][1])  # same initial grid
    xg, xa, xb = x[0], runs[0][2], runs[1][2]
    r = radial_profile(xg, nc, 5 * PC)
    cell = 10 * PC / nc
    Rg, Ra, Rb = stromgren_radius(xg, r), stromgren_radius(xa, r), stromgren_radius(xb, r)
    # the region is chosen geometrically (well inside the front), NOT by the values being
    # compared: selecting cells where both references are small biases them low
    ion = r < 0.75 * Ra
    assert ion.sum() > 1000
    noise = np.median(np.abs(xa[ion] - xb[ion]) / xa[ion])       # reference vs reference
    dev = np.median(np.abs(xg[ion] - xa[ion]) / xa[ion])         # GPU vs reference
    assert dev < 1.5 * noise + 1e-3, (dev, noise)
    # shell averages beat the per-cell noise down: 1 % agreement of the neutral-fraction profile
    edges = np.linspace(0., 0.75 * Ra, 9)
    sg, sa, sb = shell_means(xg, r, edges), shell_means(xa, r, edges), shell_means(xb, r, edges)
    tol = np.maximum(3. * np.abs(sb / sa - 1.), 0.01)
    assert (np.abs(sg / sa - 1.) < tol).all(), (sg / sa, sb / sa)
    # Stroemgren radius: within a fraction of a cell of the reference's and (no diffuse field)
    # of the analytic value (0.75 Q / (pi n^2 alpha))^(1/3) (benchmarks/stromgren.py:47-64)
    assert abs(Rg - Ra) < max(0.25 * cell, 3. * abs(Ra - Rb))
    if not diffuse:
        Rs = (0.75 * 4.26e49 / (np.pi * (1e8) ** 2 * 4e-19)) ** (1. / 3.)
        assert abs(Rg - Rs) < 1.0 * cell
    # neutral outside, ionised inside, same cells
    assert np.mean((xg > 0.5) != (xa > 0.5)) < 0.01


def test_lexington_hii20_matches_the_reference(cmib, ref, tmp_path):
    """Full physics: Planck + Verner + 14 ions + Physical diffuse field + temperature
    solve with line cooling, reduced to 32^3 / 2e5 packets / 8 iterations so the CPU
    reference finishes in seconds."""
    from cmacionize_b200 import problems
    nc, npk, nit = 32, 1000000, 8
    yml = tmp_path / "lex.yml"
    yml.write_text("""number of blocks: 2
block[0]:
  origin: [0. pc, 0. pc, 0. pc]
  sides: [6. pc, 6. pc, 6. pc]
  type: cube
  number density: 100. cm^-3
  initial temperature: 8000. K
block[1]:
  origin: [0. pc, 0. pc, 0. pc]
  sides: [6.e18 cm, 6.e18 cm, 6.e18 cm]
  type: sphere
  number density: 0. cm^-3
  initial temperature: 0. K
""")
    runs = []
    for seed in (42, 4242):
        pf = tmp_path / f"lex_{seed}.param"
        pf.write_text(f"""AbundanceModel:
  type: FixedValue
  He: 0.1
  C: 2.2e-4
  N: 4.e-5
  O: 3.3e-4
  Ne: 5.e-5
  S: 9.e-6
DensityFunction:
  type: BlockSyntax
  filename: {yml}
DiffuseReemissionHandler:
  type: Physical
SimulationBox:
  anchor: [-3. pc, -3. pc, -3. pc]
  sides: [6. pc, 6. pc, 6. pc]
  periodicity: [false, false, false]
DensityGrid:
  type: Cartesian
  number of cells: [{nc}, {nc}, {nc}]
IonizationSimulation:
  number of iterations: {nit}
  number of photons: {npk}
  random seed: {seed}
TemperatureCalculator:
  do temperature calculation: true
  PAH heating factor: 0.
PhotonSourceDistribution:
  type: SingleStar
  position: [0. pc, 0. pc, 0. pc]
  luminosity: 1.e49 s^-1
PhotonSourceSpectrum:
  type: Planck
  temperature: 20000. K
""")
        fields, _ = ref.run_paramfile(pf, nc ** 3)
        runs.append(fields)
    prob = problems.lexington(20, ncell=nc, n_packets=npk, n_iterations=nit)
    assert np.array_equal(prob.number_density, runs[0][0])   # same vacuum sphere, cell for cell
    problems.run(prob)
    n, T, x, heat = prob.ctx.download_cells()
    prob.ctx.close()
    a, b = runs
    gas = n > 0
    r = radial_profile(x[0], nc, 3 * PC)
    # geometric region: gas inside 75 % of the reference's ionisation-front radius
    Ra = stromgren_radius(np.where(gas, a[2], 0.), r)
    ion = gas & (r < 0.75 * Ra)
    assert ion.sum() > 500
    # hydrogen
    noise = np.median(np.abs(a[2][ion] - b[2][ion]) / a[2][ion])
    dev = np.median(np.abs(x[0][ion] - a[2][ion]) / a[2][ion])
    assert dev < 1.5 * noise + 1e-3, ("xH", dev, noise)
    # temperature
    noiseT = np.median(np.abs(a[1][ion] - b[1][ion]) / a[1][ion])
    devT = np.median(np.abs(T[ion] - a[1][ion]) / a[1][ion])
    assert devT < 1.5 * noiseT + 1e-3, ("T", devT, noiseT)
    assert abs(T[ion].mean() / a[1][ion].mean() - 1.) < max(3. * abs(b[1][ion].mean() / a[1][ion].mean() - 1.), 0.01)
    # every ion: volume-averaged fraction over the region.  The reference itself is not
    # reproducible run to run (OpenMP dynamic job hand-out), so the yardstick |ma - mb| is a
    # one-sample noise estimate: 4 x that plus a 5 % floor keeps the check meaningful for the
    # metals without flaking; hydrogen and temperature carry the tight checks above.
    report = {}
    for k in range(14):
        ma, mb, mg = a[2 + k][ion].mean(), b[2 + k][ion].mean(), x[k][ion].mean()
        report[k] = (float(mg), float(ma), float(mb))
    import json, os
    from pathlib import Path
    out = Path(os.environ.get("GRAFT_REPO_ROOT", ".")) / "gpurun_out"
    out.mkdir(exist_ok=True)
    (out / "parity_lexington.json").write_text(json.dumps(dict(xH_dev=float(dev), xH_noise=float(noise), T_dev=float(devT),
                                                               T_noise=float(noiseT), ion_means_gpu_refA_refB=report)))
    for k, (mg, ma, mb) in report.items():
        # + an absolute floor: fractions below ~1e-4 (Ne+, S++ at 20000 K) come from a handful of hard
        # packets; two reference runs give anything between exactly 0 and 1.5e-4 there
        tol = 4. * abs(ma - mb) + 0.05 * abs(ma) + 3e-4
        assert abs(mg - ma) < tol, (k, mg, ma, mb)
    # vacuum cells: T = 500 K, everything neutral/zero exactly as the reference leaves them
    vac = ~gas
    assert np.array_equal(T[vac], a[1][vac])
    assert np.array_equal(x[:, vac], a[2:16][:, vac])


def test_temperature_update_kernels_agree_bitwise(cmib):
    """The production temperature update (persistent warps, dynamic cell hand-out,
    update_temperature_kernel) and the one-thread-per-cell kernel run the same per-cell state
    machine on the same accumulators: every cell must come out bit-identical."""
    import os
    from cmacionize_b200 import problems
    prob = problems.lexington(20, ncell=24, n_packets=400000)
    ctx = prob.ctx
    for loop in range(5):
        problems.run_iteration(prob, loop)          # reach the temperature-solve iterations (loop > 3)
    ctx.reset_accumulators()
    ctx.update_reemission_probabilities()
    ctx.shoot(400000, seed=3, iteration=5)
    n0, T0, x0, _ = ctx.download_cells()
    results = []
    for simple in ("0", "1"):
        os.environ["CMIB_UPDATE_SIMPLE"] = simple
        ctx.upload_cells(n0, T0, x0)
        ctx.update_state(5, 0.)
        results.append(ctx.download_cells())
    os.environ.pop("CMIB_UPDATE_SIMPLE", None)
    ctx.close()
    (na, Ta, xa, ha), (nb, Tb, xb, hb) = results
    assert (Ta[n0 > 0] > 500.).mean() > 0.3          # the solve really ran
    assert np.array_equal(Ta, Tb) and np.array_equal(xa, xb, equal_nan=True) and np.array_equal(ha, hb)
