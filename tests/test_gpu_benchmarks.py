"""GPU tier: the reference's own benchmark parameter files (tests/golden/benchmarks, unmodified
copies of /root/reference/benchmarks/*.param) through the C++ host driver, on the benchmark's real
64^3 grid, against two runs of the compiled reference on the same file.

stromgren / stromgren_diffuse run at the file's full size (1e6 packets x 20 iterations); the
Lexington files say 1e8 packets x 20 iterations (~9 minutes per CPU run), so the packet count is
overridden to 1e6 x 10 iterations by appending a second `IonizationSimulation:` block (groups
merge, later keys win — YAMLDictionary.hpp:177-260).  Noise yardstick and region selection as in
test_gpu_simulation.py."""
import shutil
from pathlib import Path

import numpy as np
import pytest

from test_gpu_host_driver import host  # noqa: F401  (fixture)
from test_gpu_simulation import radial_profile, shell_means, stromgren_radius

pytestmark = pytest.mark.gpu

PC = 3.086e16
BENCH = Path(__file__).resolve().parent / "golden" / "benchmarks"
CASES = {
    "stromgren": dict(half=5., override=None, full_physics=False),
    "stromgren_diffuse": dict(half=5., override=None, full_physics=False),
    "lexingtonHII20": dict(half=3., override=(1_000_000, 10), full_physics=True),
    "lexingtonHII40": dict(half=5., override=(1_000_000, 10), full_physics=True),
}


def make_paramfile(tmp_path, name, seed):
    text = (BENCH / f"{name}.param").read_text()
    yml = BENCH / f"{name}.yml"
    if yml.exists():
        shutil.copy(yml, tmp_path / yml.name)
        text = text.replace(f"filename: {yml.name}", f"filename: {tmp_path / yml.name}")
    extra = f"\nIonizationSimulation:\n  random seed: {seed}\n  output folder: {tmp_path}\n"
    ov = CASES[name]["override"]
    if ov:
        extra += f"  number of photons: {ov[0]}\n  number of iterations: {ov[1]}\n"
    extra += f"\nTaskBasedIonizationSimulation:\n  random seed: {seed}\n  output folder: {tmp_path}\n"
    if ov:
        extra += f"  number of photons: {ov[0]}\n  number of iterations: {ov[1]}\n"
    if CASES[name]["full_physics"]:
        # the Lexington files switch the diffuse field on through `DiffuseReemissionHandler:`; the task-based driver
        # has its own switch.  The oracle is built without HDF5: both drivers get the ASCII writer (never used here)
        extra += "  diffuse field: true\nDensityGridWriter:\n  type: AsciiFile\n"
    pf = tmp_path / f"{name}_{seed}.param"
    pf.write_text(text + extra)
    return pf


@pytest.mark.parametrize("name,task_based", [(n, False) for n in CASES] + [("stromgren_diffuse", True), ("lexingtonHII20", True), ("lexingtonHII40", True)])
def test_benchmark_parameter_file(host, ref, tmp_path, name, task_based):  # noqa: F811
    """task_based: the same file through the `CMacIonize --task-based` parameter surface
    (TaskBasedIonizationSimulation: block, diffuse field switch) with that driver's packet conventions on the
    device (cmib_set_packet_conventions).  stromgren_diffuse (A_He = 0: the conventions cannot show) against the
    IonizationSimulation runs; lexingtonHII20 / HII40 (He + metals, temperature solve) against two runs of the reference's
    own TaskBasedIonizationSimulation (oracle probe cmi_ref_run_paramfile_taskbased)."""
    case = CASES[name]
    nc = 64
    runs = []
    for seed in (42, 4242):
        if task_based and case["full_physics"]:
            fields = ref.run_paramfile_taskbased(make_paramfile(tmp_path, name, seed), nc ** 3)
        else:
            fields, _ = ref.run_paramfile(make_paramfile(tmp_path, name, seed), nc ** 3)
        runs.append(fields)
    sim = host.IonizationSimulation(make_paramfile(tmp_path, name, 42), task_based=task_based)
    assert sim.number_of_photons == (1_000_000 if case["override"] is None else case["override"][0])
    assert sim.ncells == nc ** 3
    sim.initialize()
    n0 = sim.fields()[0]
    sim.run()
    n, T, x, heat = sim.fields()
    sim.close()
    a, b = runs
    assert np.array_equal(n0, a[0]) and np.array_equal(n, a[0])       # same grid, cell for cell
    gas = n > 0
    half = case["half"] * PC
    r = radial_profile(x[0], nc, half)
    cell = 2 * half / nc
    Ra = stromgren_radius(np.where(gas, a[2], 0.), r)
    Rb = stromgren_radius(np.where(gas, b[2], 0.), r)
    Rg = stromgren_radius(np.where(gas, x[0], 0.), r)
    assert abs(Rg - Ra) < max(0.25 * cell, 3. * abs(Ra - Rb)), (Rg / cell, Ra / cell, Rb / cell)
    ion = gas & (r < 0.75 * Ra)
    assert ion.sum() > 5000
    noise = np.median(np.abs(a[2][ion] - b[2][ion]) / a[2][ion])
    dev = np.median(np.abs(x[0][ion] - a[2][ion]) / a[2][ion])
    assert dev < 1.5 * noise + 1e-3, (dev, noise)
    r_in = np.sqrt(3.) * cell if not case["full_physics"] else 1.05 * 0.5 * 6.e16  # outside the vacuum sphere
    edges = np.linspace(r_in, 0.75 * Ra, 9)
    sg, sa, sb = (shell_means(f[ion], r[ion], edges) for f in (x[0], a[2], b[2]))
    tol = np.maximum(3. * np.abs(sb / sa - 1.), 0.01)
    assert (np.abs(sg / sa - 1.) < tol).all(), (sg / sa, sb / sa)    # neutral-fraction profile to 1 %
    if name == "stromgren":
        # analytic Stroemgren radius (benchmarks/stromgren.py:47-64)
        Rs = (0.75 * 4.26e49 / (np.pi * (1e8) ** 2 * 4e-19)) ** (1. / 3.)
        assert abs(Rg - Rs) < 1.0 * cell
    if case["full_physics"]:
        noiseT = np.median(np.abs(a[1][ion] - b[1][ion]) / a[1][ion])
        devT = np.median(np.abs(T[ion] - a[1][ion]) / a[1][ion])
        assert devT < 1.5 * noiseT + 1e-3, (devT, noiseT)
        assert abs(T[ion].mean() / a[1][ion].mean() - 1.) < max(3. * abs(b[1][ion].mean() / a[1][ion].mean() - 1.), 0.01)
        for k in range(14):
            ma, mb, mg = a[2 + k][ion].mean(), b[2 + k][ion].mean(), x[k][ion].mean()
            # absolute floor: fractions below ~1e-4 come from a handful of hard packets (test_gpu_simulation.py)
            assert abs(mg - ma) < 4. * abs(ma - mb) + 0.05 * abs(ma) + 3e-4, (k, mg, ma, mb)
        vac = ~gas
        assert np.array_equal(T[vac], a[1][vac]) and np.array_equal(x[:, vac], a[2:16][:, vac])
