"""CPU tier: the C-ABI library builds for sm_100a, loads without a GPU, exports
every symbol include/cmib.h declares, and refuses to compute without a device."""
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    text = (ROOT / "include" / "cmib.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(cmib_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    syms = declared_symbols()
    for s in ("cmib_create", "cmib_shoot", "cmib_update_state", "cmib_march_packets",
              "cmib_accumulator_buffer", "cmib_upload_cells", "cmib_download_cells"):
        assert s in syms
    assert len(syms) >= 30


def test_library_exports_every_declared_symbol(cmib):
    lib = cmib.capi.lib
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, missing
    assert lib.cmib_abi_version() == 1


def test_host_library_exports_the_reference_c_abi():
    """include/cmi_c_library.h = the reference's c/cmi_c_library.h: every function is exported by
    libcmih.so under the reference's name."""
    import ctypes
    text = (ROOT / "include" / "cmi_c_library.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    syms = sorted(set(re.findall(r"\b(cmi_[a-z0-9_]+)\s*\(", text)))
    assert syms == ["cmi_compute_neutral_fraction_dp", "cmi_compute_neutral_fraction_mp",
                    "cmi_compute_neutral_fraction_sp", "cmi_destroy", "cmi_init", "cmi_init_periodic_dp",
                    "cmi_init_periodic_sp"]
    lib = ctypes.CDLL(str(ROOT / "cmacionize_b200" / "libcmih.so"))
    assert not [s for s in syms if not hasattr(lib, s)]


def test_no_cpu_fallback(cmib):
    """Without a CUDA device the product must fail loudly, never compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the refusal path is exercised on the CPU box")
    with pytest.raises(cmib.CmibError, match="no CUDA device"):
        cmib.Context([0, 0, 0], [1, 1, 1], [4, 4, 4])


def test_product_does_not_use_the_oracle_or_hostcheck():
    """oracle/ and tests/hostcheck are checkers; the product may not load them."""
    pat = re.compile(r"(import\s+oracle|from\s+oracle|libcmi_ref|libhostcheck|oracle/_ref)")
    for path in (ROOT / "cmacionize_b200").rglob("*"):
        if path.suffix in (".py", ".cu", ".cuh", ".hpp", ".cpp", ".h"):
            assert not pat.search(path.read_text()), path
