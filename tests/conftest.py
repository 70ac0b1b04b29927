"""Shared fixtures.

Tiers
  -m "not gpu": oracle pinned against the reference's golden vectors, the product's
                physics headers compiled for the host and checked against the oracle,
                C-ABI library loads and exports every declared symbol, host logic,
                2-rank gloo plumbing.  No GPU needed.
  -m gpu:       parity tests proper — every check goes through the C ABI on cuda:0.
"""
from __future__ import annotations

import ctypes
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden():
    return np.load(ROOT / "tests" / "golden" / "reference_fixtures.npz")


@pytest.fixture(scope="session")
def ref():
    """The compiled reference (oracle/_ref/libcmi_ref.so)."""
    import oracle.ref as r
    if not r.available():
        if Path("/root/reference").exists():
            subprocess.check_call([sys.executable, str(ROOT / "oracle" / "build_ref.py")])
        else:
            pytest.skip("oracle/_ref/libcmi_ref.so not built and /root/reference absent")
    r.lib()
    return r


@pytest.fixture(scope="session")
def hostcheck():
    """Product physics headers compiled for the host (tests/hostcheck)."""
    src = ROOT / "tests" / "hostcheck" / "hostcheck.cpp"
    out = ROOT / "tests" / "hostcheck" / "_build" / "libhostcheck.so"
    deps = [src] + list((ROOT / "cmacionize_b200" / "csrc").glob("*"))
    if not out.exists() or any(d.stat().st_mtime > out.stat().st_mtime for d in deps):
        out.parent.mkdir(parents=True, exist_ok=True)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-fopenmp", "-ffp-contract=off", "-fPIC",
                               "-shared", "-x", "c++", str(src), "-o", str(out)])
    return ctypes.CDLL(str(out))


@pytest.fixture(scope="session")
def cmib():
    """The product binding; building the library is part of the CPU tier."""
    lib = ROOT / "cmacionize_b200" / "libcmib.so"
    if not lib.exists():
        from importlib import import_module
        sys.path.insert(0, str(ROOT))
        subprocess.check_call([sys.executable, "-m", "cmacionize_b200.build"], cwd=str(ROOT))
    import cmacionize_b200
    return cmacionize_b200


def rel_err(a, b):
    """max |a-b| / max(|a|,|b|) over finite entries; NaN patterns must agree."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    assert np.array_equal(np.isnan(a), np.isnan(b)), "NaN pattern differs"
    m = ~np.isnan(a)
    if not m.any():
        return 0.0
    a, b = a[m], b[m]
    s = np.maximum(np.abs(a), np.abs(b))
    d = np.abs(a - b)
    with np.errstate(invalid="ignore", divide="ignore"):
        r = np.where(s > 0, d / s, 0.0)
    r = np.where(np.isinf(a) & (a == b), 0.0, r)
    return float(np.max(r))
