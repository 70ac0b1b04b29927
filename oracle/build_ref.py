#!/usr/bin/env python3
"""Build the UNMODIFIED reference (CMacIonize) hot-path sources + our C harness
into ``oracle/_ref/libcmi_ref.so``.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cmacionize_b200/`` may load this
library; it is the checker for ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py``.

What this does (it does NOT run the reference's CMake build system):
  1. expands the reference's ``*.in`` templates (``Configuration.hpp`` etc.)
     into ``oracle/_ref/gen/`` with the same options the default CMake
     configure would pick on this box (OpenMP on; MPI, HDF5 off; locks, not
     lock-free; all elements; fixed abundances), cf.
     /root/reference/src/Configuration.hpp.in and src/CMakeLists.txt:23-133;
  2. copies the five atomic-data files the path reads into ``oracle/_ref/data``
     (the data-location macros are pointed at a run-time resolver in the
     harness so the .so stays relocatable);
  3. compiles the 38 sources of SURVEY.md Appendix E (+ TaskBasedIonizationSimulation.cpp, the reference's other
     driver, for the f2 parity test) *where they lie* under
     /root/reference/src with the reference's default FP semantics
     (``-std=c++11 -O3 -fopenmp``, no ``-march``, no ``-ffast-math``) plus
     ``oracle/ref_harness.cpp`` and links one shared object.

No reference source is copied into the repository; ``oracle/_ref/`` is
git-ignored (objects, generated headers, data copies, the .so).
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

HERE = Path(__file__).resolve().parent
OUT = HERE / "_ref"
REF = Path(os.environ.get("CMI_REFERENCE_ROOT", "/root/reference"))

# SURVEY.md Appendix E / reference src/CMakeLists.txt:135-172,205-215
SOURCES = """
AsciiFileDensityFunction AsciiFileDensityGridWriter AsciiFileTablePhotonSourceDistribution
ChargeTransferRates CommandLineOption CommandLineParser DeRijckeRadiativeCooling
FaucherGiguerePhotonSourceSpectrum HeliumLymanContinuumSpectrum HeliumTwoPhotonContinuumSpectrum
HydrogenLymanContinuumSpectrum InterpolatedDensityFunction IonizationStateCalculator LineCoolingData
MaskedPhotonSourceSpectrum MultiTracker ParameterFile Pegase3PhotonSourceSpectrum
PhantomSnapshotDensityFunction PhotonSource PhysicalDiffuseReemissionHandler PlanckPhotonSourceSpectrum
PopStarPhotonSourceSpectrum Signals SPHNGSnapshotDensityFunction SPHNGVoronoiGeneratorDistribution
TemperatureCalculator VernerCrossSections VernerRecombinationRates WMBasicPhotonSourceSpectrum
CartesianDensityGrid DensityGrid IonizationSimulation NewVoronoiCellConstructor NewVoronoiGrid
OldVoronoiCell OldVoronoiGrid VoronoiDensityGrid SPHArrayInterface CMILibrary TaskBasedIonizationSimulation
""".split()

DATA_FILES = ["verner_A.dat", "verner_B.dat", "verner_C.dat",
              "verner_rec_data.txt", "He2q.dat"]

# same defaults the stock configure chooses here (CMakeLists.txt:140-190)
DEFINES_ON = {"HAVE_OPENMP", "HAVE_POSIX", "HAVE_ATOMIC"}

CXXFLAGS = ["-std=c++11", "-O3", "-fopenmp", "-fPIC", "-include", "cstdint", "-w"]


def expand_template(src: Path, dst: Path, subst: dict[str, str]) -> None:
    text = src.read_text()

    def cmakedefine(m: re.Match) -> str:
        name = m.group(1)
        return f"#define {name}" if name in DEFINES_ON else f"/* #undef {name} */"

    text = re.sub(r"#cmakedefine\s+(\w+)", cmakedefine, text)
    text = re.sub(r"@(\w+)@", lambda m: subst.get(m.group(1), ""), text)
    dst.write_text(text)


def generate_headers(gen: Path) -> None:
    gen.mkdir(parents=True, exist_ok=True)
    src = REF / "src"
    resolver = "(cmi_ref_data_file(\"%s\"))"
    subst = {
        "MAX_NUM_THREADS": "512",
        "CONFIGURATION_OPTIONS_NUMBER": "1",
        "CONFIGURATION_OPTIONS_KEYS": '"oracle_build"',
        "CONFIGURATION_OPTIONS_VALUES": '"oracle/build_ref.py"',
        "GIT_BUILD_STRING": "oracle-build",
        "COMPILATION_TIME_DAY": time.strftime("%d"),
        "COMPILATION_TIME_MONTH": time.strftime("%m"),
        "COMPILATION_TIME_YEAR": time.strftime("%Y"),
        "COMPILATION_TIME_HOUR": time.strftime("%H"),
        "COMPILATION_TIME_MINUTES": time.strftime("%M"),
        "COMPILATION_TIME_SECONDS": time.strftime("%S"),
        "COMPILER_NAME": "GNU", "COMPILER_VERSION": "g++",
        "OS_NAME": "Linux", "OS_KERNEL_NAME": "Linux", "OS_KERNEL_RELEASE": "",
        "OS_KERNEL_VERSION": "", "OS_HARDWARE_NAME": "x86_64", "OS_HOST_NAME": "oracle",
        # unused spectra: plain (dangling) literal paths keep literal concatenation legal
        "FAUCHERGIGUEREDATALOCATION": "/nonexistent/fg/",
        "DERIJCKEDATALOCATION": "/nonexistent/derijcke/",
        "WMBASICDATALOCATION": "/nonexistent/wmbasic/",
        "PEGASE3DATALOCATION": "/nonexistent/pegase3/",
        "POPSTARDATALOCATION": "/nonexistent/popstar/",
        "CASTELLIKURUCZDATALOCATION": "/nonexistent/ck.hdf5",
    }
    for tmpl in sorted(src.glob("*.in")):
        expand_template(tmpl, gen / tmpl.name[:-3], subst)
    # the three data locations the hot path really uses go through a resolver so
    # the library finds oracle/_ref/data next to itself wherever the repo lives
    decl = "#include <string>\nextern \"C++\" std::string cmi_ref_data_file(const char *name);\n"
    (gen / "VernerCrossSectionsDataLocation.hpp").write_text(
        "#ifndef VERNERCROSSSECTIONSDATALOCATION_HPP\n#define VERNERCROSSSECTIONSDATALOCATION_HPP\n"
        + decl
        + "#define VERNERCROSSSECTIONSDATALOCATION_A " + resolver % "verner_A.dat" + "\n"
        + "#define VERNERCROSSSECTIONSDATALOCATION_B " + resolver % "verner_B.dat" + "\n"
        + "#define VERNERCROSSSECTIONSDATALOCATION_C " + resolver % "verner_C.dat" + "\n#endif\n")
    (gen / "VernerRecombinationRatesDataLocation.hpp").write_text(
        "#ifndef VERNERRECOMBINATIONRATESDATALOCATION_HPP\n#define VERNERRECOMBINATIONRATESDATALOCATION_HPP\n"
        + decl
        + "#define VERNERRECOMBINATIONRATESDATALOCATION " + resolver % "verner_rec_data.txt" + "\n#endif\n")
    (gen / "FaucherGiguereDataLocation.hpp").write_text(
        "#ifndef FAUCHERGIGUEREDATALOCATION_HPP\n#define FAUCHERGIGUEREDATALOCATION_HPP\n"
        "#include <string>\nextern \"C++\" std::string cmi_ref_data_dir(const char *sub, const char *probe_file);\n"
        "#define FAUCHERGIGUEREDATALOCATION (cmi_ref_data_dir(\"fg_uvb_dec11\", \"fg_uvb_dec11_z_0.0.dat\"))\n#endif\n")
    (gen / "HeliumTwoPhotonContinuumDataLocation.hpp").write_text(
        "#ifndef HELIUMTWOPHOTONCONTINUUMDATALOCATION_HPP\n#define HELIUMTWOPHOTONCONTINUUMDATALOCATION_HPP\n"
        + decl
        + "#define HELIUMTWOPHOTONCONTINUUMDATALOCATION " + resolver % "He2q.dat" + "\n#endif\n")


def run(cmd: list[str]) -> None:
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        raise SystemExit(f"oracle build failed: {cmd[-1]}")


def newer(a: Path, b: Path) -> bool:
    return (not b.exists()) or a.stat().st_mtime > b.stat().st_mtime


def build(force: bool = False, jobs: int | None = None) -> Path:
    lib = OUT / "libcmi_ref.so"
    if not REF.exists():
        if lib.exists():
            return lib  # GPU box: prebuilt library travels with the snapshot
        raise SystemExit(f"{REF} not present and no prebuilt {lib}")
    gen, obj, data = OUT / "gen", OUT / "obj", OUT / "data"
    obj.mkdir(parents=True, exist_ok=True)
    data.mkdir(parents=True, exist_ok=True)
    if force or not (gen / "Configuration.hpp").exists():
        generate_headers(gen)
    for f in DATA_FILES:
        if force or not (data / f).exists():
            shutil.copyfile(REF / "data" / f, data / f)
    if force or not (data / "fg_uvb_dec11" / "fg_uvb_dec11_z_0.0.dat").exists():
        import tarfile
        with tarfile.open(REF / "data" / "fg_uvb_dec11.tar.gz") as tar:  # src/CMakeLists.txt unpacks it the same way
            tar.extractall(data)

    inc = [f"-I{gen}", f"-I{REF / 'src'}"]
    work: list[tuple[Path, Path]] = []
    for s in SOURCES:
        work.append((REF / "src" / f"{s}.cpp", obj / f"{s}.o"))
    work.append((gen / "CompilerInfo.cpp", obj / "CompilerInfo.o"))
    work.append((gen / "ConfigurationInfo.cpp", obj / "ConfigurationInfo.o"))
    harness = HERE / "ref_harness.cpp"
    work.append((harness, obj / "ref_harness.o"))

    todo = [(s, o) for s, o in work if force or newer(s, o)]
    jobs = jobs or os.cpu_count() or 4
    with ThreadPoolExecutor(max_workers=jobs) as ex:
        list(ex.map(lambda so: run(["g++", *CXXFLAGS, *inc, "-c", str(so[0]), "-o", str(so[1])]), todo))
    if todo or not lib.exists():
        run(["g++", "-shared", "-fopenmp", "-o", str(lib), *[str(o) for _, o in work], "-ldl"])
    return lib


if __name__ == "__main__":
    t0 = time.time()
    p = build(force="--force" in sys.argv)
    print(f"built {p} in {time.time() - t0:.1f} s")
