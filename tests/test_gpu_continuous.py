"""GPU tier: continuous photon sources (src/IsotropicContinuousPhotonSource.hpp,
src/PlanarContinuousPhotonSource.hpp, src/DistantStarContinuousPhotonSource.hpp,
src/ExtendedDiscContinuousPhotonSource.hpp, src/SpiralGalaxyContinuousPhotonSource.hpp,
PhotonSource.cpp:100-131, 208-249) end to end through the C++ host
driver — parameter files with an external radiation field alone, a star + external field, and an
emitting sheet in the mid-plane (tests/golden/continuous/), against two runs of the compiled reference on
the same file (seeds 42 / 4242).

The geometry is not spherical, so cells are compared directly: the per-cell deviation from reference
run A must not exceed the deviation between the two reference runs (Monte Carlo noise), the ionised
volume and the mean neutral fraction per depth-from-the-nearest-face shell must agree."""
from pathlib import Path

import numpy as np
import pytest

from test_gpu_host_driver import host  # noqa: F401  (fixture)

pytestmark = pytest.mark.gpu

FILES = Path(__file__).resolve().parent / "golden" / "continuous"


def make_paramfile(tmp_path, name, seed):
    text = (FILES / f"{name}.param").read_text()
    pf = tmp_path / f"{name}_{seed}.param"
    pf.write_text(text + f"\nIonizationSimulation:\n  random seed: {seed}\n  output folder: {tmp_path}\n")
    return pf


@pytest.mark.parametrize("name", ["external_field", "star_plus_external_field", "planar_sheet", "distant_star",
                                  "extended_disc", "spiral_galaxy"])
def test_external_radiation_field(host, ref, tmp_path, name):  # noqa: F811
    nc = 32
    runs = [ref.run_paramfile(make_paramfile(tmp_path, name, seed), nc ** 3)[0] for seed in (42, 4242)]
    sim = host.IonizationSimulation(make_paramfile(tmp_path, name, 42))
    sim.initialize()
    sim.run()
    n, T, x, heat = sim.fields()
    sim.close()
    a, b = runs[0][2], runs[1][2]
    g = x[0]
    # ionised volume (cells)
    va, vb, vg = (a < 0.5).sum(), (b < 0.5).sum(), (g < 0.5).sum()
    assert 0.1 * nc ** 3 < va < 0.9 * nc ** 3
    assert abs(vg - va) <= max(3 * abs(va - vb), 0.005 * va), (vg, va, vb)
    # per-cell deviation in the ionised region, against the reference's own noise
    ion = (a < 0.1) & (b < 0.1)
    noise = np.median(np.abs(a[ion] - b[ion]) / a[ion])
    dev = np.median(np.abs(g[ion] - a[ion]) / a[ion])
    assert dev < 1.5 * noise + 1e-3, (dev, noise)
    # mean neutral fraction per shell of equal depth below the nearest face
    i = np.arange(nc)
    depth1 = np.minimum(i, nc - 1 - i)
    if name in ("planar_sheet", "extended_disc", "spiral_galaxy"):   # distance from the sheet / the mid-plane of the disc instead
        depth = np.meshgrid(i, i, np.abs(i - (nc - 1) / 2.).astype(int), indexing="ij")[2].ravel()
    elif name == "distant_star":  # depth below the two lit faces (x low, y low)
        depth = np.minimum(*np.meshgrid(i, i, i, indexing="ij")[:2]).ravel() // 2
    else:
        depth = np.minimum.reduce(np.meshgrid(depth1, depth1, depth1, indexing="ij")).ravel()
    for d in range(nc // 2):
        sel = depth == d
        ma, mb, mg = a[sel].mean(), b[sel].mean(), g[sel].mean()
        assert abs(mg - ma) <= 4 * abs(ma - mb) + 0.02 * ma + 1e-6, (d, mg, ma, mb)
