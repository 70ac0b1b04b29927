"""Seeded input generators shared by the CPU-tier and GPU-tier parity tests."""
from __future__ import annotations

import numpy as np

PC = 3.086e16
ABUNDANCES = np.array([0.1, 2.2e-4, 4e-5, 3.3e-4, 5e-5, 9e-6])  # He C N O Ne S (Lexington)

# (anchor, sides, ncell, periodic, density scale, vacuum?, source at box centre?)
MARCH_GRIDS = {
    "unit16": dict(anchor=[0, 0, 0], sides=[1, 1, 1], ncell=[16, 16, 16], periodic=[0, 0, 0],
                   dens=1e12, vac=False, corner=False),
    "stromgren64_corner": dict(anchor=[-5 * PC] * 3, sides=[10 * PC] * 3, ncell=[64, 64, 64],
                               periodic=[0, 0, 0], dens=1., vac=False, corner=True),
    "noncubic_periodic_xz": dict(anchor=[-1, -2, -3], sides=[2, 5, 3], ncell=[8, 20, 12],
                                 periodic=[1, 0, 1], dens=1e14, vac=False, corner=False),
    "tiny_fully_periodic": dict(anchor=[-1, -2, -3], sides=[2, 5, 3], ncell=[7, 5, 3],
                                periodic=[1, 1, 1], dens=1e14, vac=False, corner=False),
    "vacuum_holes": dict(anchor=[-1, -2, -3], sides=[2, 5, 3], ncell=[9, 5, 11],
                         periodic=[0, 0, 0], dens=1e13, vac=True, corner=False),
    "single_cell": dict(anchor=[0, 0, 0], sides=[1, 2, 3], ncell=[1, 1, 1], periodic=[0, 0, 0],
                        dens=1e13, vac=False, corner=False),
}

# the full grid size of BASELINE.json configs[4] (tests/test_gpu_parity256.py only: 16.8 M cells); packets start
# anywhere; mean optical depth per cell ~0.01, so the sampled depths (x 0.01, 1, 100) give early absorptions,
# walks of a few hundred cells and escapes
BIG_MARCH_GRIDS = {
    "clumpy256": dict(anchor=[-5 * PC] * 3, sides=[10 * PC] * 3, ncell=[256, 256, 256], periodic=[0, 0, 0],
                      dens=1e-2, vac=False, corner=False),
}


def march_case(name: str, npackets: int, seed: int = 7):
    """Random cells + explicit packets for CartesianDensityGrid::interact parity."""
    g = MARCH_GRIDS.get(name) or BIG_MARCH_GRIDS[name]
    rng = np.random.default_rng(seed)
    anchor = np.array(g["anchor"], float)
    sides = np.array(g["sides"], float)
    ncell = np.array(g["ncell"], np.int32)
    periodic = np.array(g["periodic"], np.int32)
    nc = int(np.prod(ncell))
    n = g["dens"] * np.exp(rng.uniform(np.log(1e6), np.log(1e9), nc))
    xH = np.exp(rng.uniform(np.log(1e-6), 0, nc))
    xHe = np.exp(rng.uniform(np.log(1e-6), 0, nc))
    if g["vac"]:
        n[rng.uniform(size=nc) < 0.2] = 0.
    if g["corner"]:
        pos = np.tile(anchor + sides * 0.5, (npackets, 1))  # exactly on a cell corner
    else:
        pos = anchor + sides * rng.uniform(0, 1, (npackets, 3))
    ct = rng.uniform(-1, 1, npackets)
    st = np.sqrt(1 - ct * ct)
    ph = rng.uniform(0, 2 * np.pi, npackets)
    d = np.stack([st * np.cos(ph), st * np.sin(ph), ct], 1)
    k = min(5, npackets)
    # axis-aligned, face-diagonal and body-diagonal rays: zero components and exact ties
    d[:k] = np.array([[1, 0, 0], [0, -1, 0], [0, 0, 1], [np.sqrt(.5), np.sqrt(.5), 0],
                      [1 / np.sqrt(3)] * 3])[:k]
    sig = np.abs(rng.normal(0, 1, (npackets, 14))) * 1e-22
    sig[:, 0] = 6.3e-22 * rng.uniform(0.1, 1, npackets)
    sig[rng.uniform(size=(npackets, 14)) < 0.3] = 0.
    sig[:, 0] = np.maximum(sig[:, 0], 1e-23)
    she = 0.1 * sig[:, 1]
    nu = 3.3e15 * rng.uniform(1, 4, npackets)
    w = rng.choice([1.0, 0.5, 2.0], npackets)
    tau = -np.log(rng.uniform(size=npackets)) * rng.choice([0.01, 1, 100], npackets)
    return dict(anchor=anchor, sides=sides, ncell=ncell, periodic=periodic, n=n, xH=xH, xHe=xHe,
                pos=np.ascontiguousarray(pos), dir=np.ascontiguousarray(d), sigma=sig,
                sigma_He_corr=she, nu=nu, weight=w, tau=tau)


def state_cells(golden, reps: int = 20, seed: int = 5):
    """Realistic (J, heat, n, T) cells: the reference's tbal fixture rows, jittered,
    plus the edge cases the reference special-cases (J = 0, vacuum, J_He = 0, J_H = 0)."""
    tb = golden["tbal"]
    tb = tb[tb[:, 16] <= 30000.]
    n0 = tb.shape[0]
    rng = np.random.default_rng(seed)
    J = np.tile(tb[:, :14].T, (1, reps)) * np.exp(rng.normal(0, 1.5, (14, n0 * reps)))
    heat = np.tile(tb[:, 14:16].T * 1e-7, (1, reps)) * np.exp(rng.normal(0, 1.0, (2, n0 * reps)))
    nd = np.tile(tb[:, 17] * 1e6, reps) * np.exp(rng.normal(0, 1, n0 * reps))
    T = np.tile(tb[:, 16], reps) * rng.uniform(0.3, 1.5, n0 * reps)
    J[:, :30] = 0
    nd[30:40] = 0
    J[1, 40:60] = 0
    J[0, 60:80] = 0
    return (np.ascontiguousarray(J), np.ascontiguousarray(heat), np.ascontiguousarray(nd),
            np.ascontiguousarray(T))


def wall_intersection_scenarios():
    """The nine get_wall_intersection scenarios of the reference's own unit test
    (/root/reference/test/testCartesianDensityGrid.cpp:310-465): unit box, 16^3 cells, a packet at
    (0.51, 0.51, 0.51) in cell (8, 8, 8); six axis-aligned rays, a generic ray, an edge hit and a
    corner hit.  Returned: directions, expected next_index, expected intersection, expected ds.
    Here they are run through the walk itself: the start cell is almost transparent and every other
    cell opaque, so the packet ends (to ~1e-23) on the first wall, in the cell next_index points to,
    and J_H of the start cell / (w sigma_H) is ds."""
    o = 0.51
    lo, hi = 0.5, 0.5 + 1. / 16.
    s3 = np.array([1., 2., -3.]) / np.sqrt(14.)
    rows = [
        ([1., 0., 0.], (1, 0, 0), (hi, o, o), 1. / 16. - 0.01),
        ([-1., 0., 0.], (-1, 0, 0), (lo, o, o), 0.01),
        ([0., 1., 0.], (0, 1, 0), (o, hi, o), 1. / 16. - 0.01),
        ([0., -1., 0.], (0, -1, 0), (o, lo, o), 0.01),
        ([0., 0., 1.], (0, 0, 1), (o, o, hi), 1. / 16. - 0.01),
        ([0., 0., -1.], (0, 0, -1), (o, o, lo), 0.01),
        (list(s3), (0, 0, -1), (o + 0.01 / 3., o + 0.02 / 3., lo), 0.0124722),
        (list(np.array([0., 1., 1.]) / np.sqrt(2.)), (0, 1, 1), (o, hi, hi), 0.0742462),
        (list(np.array([1., 1., 1.]) / np.sqrt(3.)), (1, 1, 1), (hi, hi, hi), 0.0909327),
    ]
    d = np.array([r[0] for r in rows])
    nxt = np.array([r[1] for r in rows])
    hit = np.array([r[2] for r in rows])
    ds = np.array([r[3] for r in rows])
    nc = 16 ** 3
    start = 8 * 256 + 8 * 16 + 8   # test/testCartesianDensityGrid.cpp:65-68
    n = np.full(nc, 1e45)   # n sigma = 1e23 per unit length: absorbed ~1e-23 behind the wall
    n[start] = 1.
    npk = len(rows)
    sig = np.zeros((npk, 14)); sig[:, 0] = 1e-22
    return dict(anchor=np.zeros(3), sides=np.ones(3), ncell=np.array([16] * 3, np.int32),
                periodic=np.zeros(3, np.int32), n=n, xH=np.ones(nc), xHe=np.zeros(nc),
                pos=np.full((npk, 3), o), dir=np.ascontiguousarray(d), sigma=sig, sigma_He_corr=np.zeros(npk),
                nu=np.full(npk, 3.3e15), weight=np.ones(npk), tau=np.full(npk, 1.),
                start=start, next_index=nxt, intersection=hit, ds=ds)


def check_wall_intersection_scenarios(c, fpos, fcell, nsteps, trace, J):
    start = c["start"]
    expect_next = start + c["next_index"][:, 0] * 256 + c["next_index"][:, 1] * 16 + c["next_index"][:, 2]
    assert (nsteps == 2).all()
    assert (trace[:, 0] == start).all()
    assert np.array_equal(trace[:, 1], expect_next)       # next_index, incl. the edge and corner hits
    assert np.array_equal(fcell, expect_next)
    assert np.abs(fpos - c["intersection"]).max() < 1e-15  # the wall point (absorbed ~1e-18 behind it)
    # ds of the nine crossings: all nine packets add ds * w * sigma_H to the start cell
    assert abs(J[0][start] / 1e-22 - c["ds"].sum()) < 1e-4 * c["ds"].sum()   # the reference's 1e-4 tolerance
