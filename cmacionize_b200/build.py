#!/usr/bin/env python3
"""Build cmacionize_b200/libcmib.so (the C-ABI library) with nvcc for sm_100a.

In-tree build so that the .so travels with the repository snapshot to the GPU
box.  Cross-compiles without a GPU.  `python -m cmacionize_b200.build [--force]`.
"""
from __future__ import annotations

import subprocess
import sys
import time
from pathlib import Path

HERE = Path(__file__).resolve().parent
CSRC = HERE / "csrc"
LIB = HERE / "libcmib.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    # the reference is built without FMA contraction; keep every FP64
    # product/sum separately rounded so device results differ from it by libm ulps only
    "-fmad=false",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-Xptxas", "-v",
    "-shared",
]


def sources_newer_than(target: Path) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    deps = list(CSRC.glob("*")) + [HERE.parent / "include" / "cmib.h"]
    return any(p.stat().st_mtime > t for p in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not sources_newer_than(LIB):
        return LIB
    cmd = ["nvcc", *NVCC_FLAGS, "-o", str(LIB), str(CSRC / "cmib_api.cu")]
    t0 = time.time()
    r = subprocess.run(cmd, capture_output=True, text=True)
    (HERE / "build.log").write_text(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise SystemExit("nvcc failed")
    if verbose:
        print(r.stderr)
    print(f"built {LIB} in {time.time() - t0:.1f} s")
    return LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
