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


HOST = HERE / "host"
HOST_LIB = HERE / "libcmih.so"
HOST_BIN = HERE / "bin" / "CMacIonizeB200"


def build_host(force: bool = False) -> Path:
    """C++ host layer (parameter file, plugin classes, IonizationSimulation driver) as a shared
    library over libcmib.so + the command line program.  Plain g++: no CUDA in this layer."""
    deps = list(HOST.glob("*")) + [HERE.parent / "include" / "cmib.h"]
    stale = (not HOST_LIB.exists() or not HOST_BIN.exists()
             or any(p.stat().st_mtime > min(HOST_LIB.stat().st_mtime, HOST_BIN.stat().st_mtime) for p in deps))
    if not force and not stale:
        return HOST_LIB
    HOST_BIN.parent.mkdir(exist_ok=True)
    common = ["g++", "-std=c++17", "-O2", "-Wall", "-Wno-unknown-pragmas", "-ffp-contract=off", "-fPIC", "-pthread"]
    link = [f"-L{HERE}", "-lcmib", "-ldl", "-lz", "-Wl,-rpath,$ORIGIN"]  # NCCL is dlopen-ed on demand
    subprocess.check_call(common + ["-shared", str(HOST / "host_api.cpp"), "-o", str(HOST_LIB)] + link)
    link_bin = [f"-L{HERE}", "-lcmib", "-ldl", "-lz", "-Wl,-rpath,$ORIGIN/.."]
    subprocess.check_call(common + [str(HOST / "CMacIonizeB200.cpp"), "-o", str(HOST_BIN)] + link_bin)
    print(f"built {HOST_LIB} and {HOST_BIN}")
    return HOST_LIB


if __name__ == "__main__":
    build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    build_host(force="--force" in sys.argv)
