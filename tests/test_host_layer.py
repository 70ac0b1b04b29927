"""CPU tier: the C++ host layer (cmacionize_b200/host) reads the reference's parameter files
exactly like the reference does — same keys, defaults, unit conversions (bit for bit), same
used-values dump, same initial grids from the DensityFunction plugins.  No GPU needed: only the
parts of the host layer that do not create a device context are exercised here; the full
IonizationSimulation driver is covered by the GPU tier (tests/test_gpu_host_driver.py)."""
import numpy as np
import pytest

PC = 3.086e16

LEX_YML = """number of blocks: 2
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
"""

PARAM = """# a comment line
AbundanceModel:
  type: FixedValue
  He: 0.1
  C: 2.2e-4   # trailing comment
DensityFunction:
  type: BlockSyntax
  filename: {yml}
SimulationBox:
  anchor: [-3. pc, -3. pc, -3. pc]
  sides: [6. pc, 6. pc, 6. pc]
  periodicity: [false, false, false]
DensityGrid:
  type: Cartesian
  number of cells: [12, 12, 12]
IonizationSimulation:
  number of iterations: 3
  number of photons: 1e5
  random seed: 42
TemperatureCalculator:
  do temperature calculation: true
  cosmic ray heating scale length: 2. kpc
PhotonSourceDistribution:
  type: SingleStar
  position: [0. pc, 0. pc, 0. pc]
  luminosity: 1.e49 s^-1
PhotonSourceSpectrum:
  type: Planck
  temperature: 20000. K
CrossSections:
  type: FixedValue
  hydrogen_0: 6.3e-18 cm^2
RecombinationRates:
  type: FixedValue
  hydrogen_1: 4.e-13 cm^3 s^-1
Nested:
  Deeper:
    key: 13.6 eV
    wavelength: 912. angstrom
"""

QUERIES = [
    ("s", "AbundanceModel:type", "FixedValue"), ("d", "AbundanceModel:He", "0."), ("d", "AbundanceModel:C", "0."),
    ("d", "AbundanceModel:N", "0."), ("s", "DensityGrid:type", "Cartesian"),
    ("L", "TemperatureCalculator:cosmic ray heating scale length", "1.33333 kpc"),
    ("T", "TemperatureCalculator:minimum ionized temperature", "4000. K"),
    ("F", "PhotonSourceDistribution:luminosity", "4.26e49 s^-1"), ("T", "PhotonSourceSpectrum:temperature", "4.e4 K"),
    ("A", "CrossSections:hydrogen_0", "6.3e-18 cm^2"), ("A", "CrossSections:helium_0", "0. m^2"),
    ("R", "RecombinationRates:hydrogen_1", "2.7e-13 cm^3 s^-1"), ("R", "RecombinationRates:helium_1", "0. m^3 s^-1"),
    ("F", "Nested:Deeper:key", "1. Hz"), ("F", "Nested:Deeper:wavelength", "1. Hz"),
    ("F", "PhotonSourceSpectrum:frequency", "13.6 eV"), ("N", "DensityFunction:density", "100. cm^-3"),
    ("d", "IonizationSimulation:number of photons", "1e5"),
    ("d", "TemperatureCalculator:epsilon convergence", "1e-3"),
]
KIND = {"L": "length", "T": "temperature", "F": "frequency", "A": "surface_area", "R": "reaction_rate",
        "N": "number_density"}


@pytest.fixture(scope="module")
def host(cmib):
    import subprocess, sys
    from conftest import ROOT
    subprocess.check_call([sys.executable, "-c", "from cmacionize_b200 import build as b; b.build_host()"],
                          cwd=str(ROOT))
    from cmacionize_b200 import host as h
    return h


def test_unit_conversions_bitexact(host, ref):
    cases = [(13.6, "eV", "Hz"), (24.6, "eV", "Hz"), (19.8, "eV", "Hz"), (912., "angstrom", "Hz"),
             (3.288465385e15, "Hz", "eV"), (1., "pc", "m"), (10., "pc", "cm"), (1.33333, "kpc", "m"),
             (100., "cm^-3", "m^-3"), (6.3e-18, "cm^2", "m^2"), (4.e-13, "cm^3 s^-1", "m^3 s^-1"),
             (2.7e-13, "cm^3 s^-1", "m^3 s^-1"), (1., "Myr", "s"), (1., "g cm^-3", "kg m^-3"),
             (5., "km s^-1", "m s^-1"), (1., "Msol", "g"), (3., "K kg^3 s^-1m ", "K kg^3 s^-1 m"),
             (1.e49, "s^-1", "Hz"), (1., "erg", "J"), (2., "au", "km")]
    for v, a, b in cases:
        assert host.convert(v, a, b) == ref.convert(v, a, b), (v, a, b)
    with pytest.raises(host.HostError, match="Unknown unit"):
        host.convert(1., "parsec", "m")


def test_parameter_file_same_values_and_used_values_dump(host, ref, tmp_path):
    yml = tmp_path / "blocks.yml"
    yml.write_text(LEX_YML)
    pf = tmp_path / "test.param"
    pf.write_text(PARAM.format(yml=yml))
    want, dump = ref.paramfile_query(pf, QUERIES)
    p = host.ParameterFile(pf)
    got = []
    for kind, key, default in QUERIES:
        if kind == "s":
            got.append(p.get_string(key, default))
        elif kind == "d":
            got.append(repr(p.get_double(key, float(default))))
        else:
            got.append(repr(p.get_physical(KIND[kind], key, default)))
    for (kind, key, _), g, w in zip(QUERIES, got, want):
        if kind == "s":
            assert g == w, key
        else:
            assert float(g) == float(w), (key, g, w)   # bit-identical doubles
    assert p.used_values() == dump
    # a mandatory key that is absent -> the reference's "Parameter ... not found!" error
    bad = tmp_path / "bad.param"
    bad.write_text("DensityFunction:\n  type: BlockSyntax\n")
    with pytest.raises(host.HostError, match="not found"):
        host.ParameterFile(bad).density_function(np.zeros((1, 3)))
    # unknown plugin type -> the factory's error
    bad.write_text("DensityFunction:\n  type: Fractal\n")
    with pytest.raises(host.HostError, match="Unknown DensityFunction type"):
        host.ParameterFile(bad).density_function(np.zeros((1, 3)))


ANALYTIC_PROFILES = {
    # type -> (DensityFunction block, box half size in pc); defaults and non-default values
    # BondiProfile: Lambert W on both branches (inside / outside the Bondi radius), with and without the ionised core
    "bondi_defaults": ("  type: BondiProfile\n", 0.02),
    # the ionised profile of the reference's testBondiProfile.cpp:56 (30 au ionisation radius, pressure contrast 32)
    "bondi_ionised_core": ("  type: BondiProfile\n  central mass: 18. Msol\n  Bondi density: 1.e-16 kg m^-3\n"
                           "  sound speed: 2.031 km s^-1\n  ionisation radius: 30. au\n  pressure contrast: 32.\n"
                           "  center: [3. au, -5. au, 1. au]\n  neutral fraction: 0.9\n", 5.e-4),
    "cored_dm_defaults": ("  type: CoredDMProfile\n", 500.),
    "cored_dm": ("  type: CoredDMProfile\n  core radius: 120. pc\n  maximum circular velocity: 15. km s^-1\n"
                 "  central density: 2.e-21 g cm^-3\n  temperature: 8000. K\n  neutral fraction: 0.3\n"
                 "  polytropic index: 1.4\n", 500.),
    "disc_ic_defaults": ("  type: DiscIC\n", 5.e-4),
    "disc_ic_hot": ("  type: DiscIC\n  mass: 12. Msol\n  temperature: 2.e4 K\n  Bondi density: 1.e-18 g cm^-3\n"
                    "  density power: 1.2\n", 5.e-4),
    "disc_patch_defaults": ("  type: DiscPatch\n", 1000.),
    "disc_patch": ("  type: DiscPatch\n  disc z: 50. pc\n  surface density: 12. Msol pc^-2\n  scale height: 350. pc\n"
                   "  gas fraction: 0.25\n  temperature: 6000. K\n  neutral fraction: 0.5\n", 1000.),
    "spiral_galaxy_defaults": ("  type: SpiralGalaxy\n", 16000.),
    "spiral_galaxy": ("  type: SpiralGalaxy\n  scale length ISM: 4. kpc\n  scale height ISM: 0.5 kpc\n"
                      "  central density: 3. cm^-3\n", 16000.),
}


@pytest.mark.parametrize("kind", ["homogeneous_defaults", "lexington_blocks", "ascii_file", "interpolated_1d",
                                  *ANALYTIC_PROFILES])
def test_density_function_gives_the_reference_grid(host, ref, tmp_path, kind):
    nc = 12
    if kind in ANALYTIC_PROFILES:
        # closed-form profiles: CoredDMProfile, DiscIC, DiscPatch, SpiralGalaxy (exp / log / atan / pow / cosh
        # per cell midpoint; the same libm on both sides, so the grids must be the same doubles)
        block, half_pc = ANALYTIC_PROFILES[kind]
        text = (f"SimulationBox:\n  anchor: [-{half_pc!r} pc, -{half_pc!r} pc, -{half_pc!r} pc]\n"
                f"  sides: [{2 * half_pc!r} pc, {2 * half_pc!r} pc, {2 * half_pc!r} pc]\n"
                "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n"
                f"  number of cells: [{nc}, {nc}, {nc}]\nDensityFunction:\n{block}"
                "PhotonSourceSpectrum:\n  type: Monochromatic\n")
        half = None
    elif kind.startswith("interpolated"):
        # InterpolatedDensityFunction: a z profile (the reference's own test file format,
        # test/test_interpolated_density.txt) sampled by the 12^3 grid
        rng = np.random.default_rng(9)
        # (one non-trivial axis: the reference's row bookkeeping never rewinds an axis index,
        # InterpolatedDensityFunction.cpp:213-247, so 2-D / 3-D tables run out of bounds there)
        zs = np.linspace(-5., 5., 23)
        head = ("---\nnum_x: 0\nxmin: -5. pc\nxmax: 5. pc\nnum_y: 0\nymin: -5. pc\nymax: 5. pc\n"
                "num_z: 23\nzmin: -5. pc\nzmax: 5. pc\nnum_column: 2\ncolumn_0_variable: z\ncolumn_0_unit: pc\n"
                "column_1_variable: number density\ncolumn_1_unit: cm^-3\n---\n")
        rows = [f"{float(z)!r} {float(rng.uniform(1., 300.))!r}" for z in zs]
        dfile = tmp_path / "profile.txt"
        dfile.write_text(head + "\n".join(rows) + "\n")
        text = ("SimulationBox:\n  anchor: [-5. pc, -5. pc, -5. pc]\n  sides: [10. pc, 10. pc, 10. pc]\n"
                "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n"
                f"  number of cells: [{nc}, {nc}, {nc}]\nDensityFunction:\n  type: Interpolated\n  filename: {dfile}\n"
                "  temperature: 6000. K\nPhotonSourceSpectrum:\n  type: Monochromatic\n")
        half = 5 * PC
    elif kind == "ascii_file":
        # AsciiFileDensityFunction: a 6 x 4 x 3 table of its own (cell centres in pc, densities in cm^-3)
        # sampled by the 12^3 simulation grid
        rng = np.random.default_rng(4)
        nf = (6, 4, 3)
        rows = ["# x y z density"]
        for i in range(nf[0]):
            for j in range(nf[1]):
                for k in range(nf[2]):
                    c = [-5. + 10. * (q + 0.5) / n for q, n in zip((i, j, k), nf)]
                    rows.append(f"{c[0]!r} {c[1]!r} {c[2]!r} {float(rng.uniform(1., 200.))!r}")
        dfile = tmp_path / "density.txt"
        dfile.write_text("\n".join(rows) + "\n")
        text = ("SimulationBox:\n  anchor: [-5. pc, -5. pc, -5. pc]\n  sides: [10. pc, 10. pc, 10. pc]\n"
                "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n"
                f"  number of cells: [{nc}, {nc}, {nc}]\nDensityFunction:\n  type: AsciiFile\n  filename: {dfile}\n"
                "  number of cells: [6, 4, 3]\n  length unit: 1. pc\n  density unit: 1. cm^-3\n  temperature: 7000. K\n"
                "PhotonSourceSpectrum:\n  type: Monochromatic\n")
        half = 5 * PC
    elif kind == "homogeneous_defaults":
        text = ("SimulationBox:\n  anchor: [-5. pc, -5. pc, -5. pc]\n  sides: [10. pc, 10. pc, 10. pc]\n"
                "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n"
                f"  number of cells: [{nc}, {nc}, {nc}]\nDensityFunction:\n  type: Homogeneous\n"
                "  neutral fraction H: 1.e-4\nPhotonSourceSpectrum:\n  type: Monochromatic\n")
        half = 5 * PC
    else:
        yml = tmp_path / "blocks.yml"
        yml.write_text(LEX_YML)
        text = PARAM.format(yml=yml)
        half = 3 * PC
    pf = tmp_path / "grid.param"
    pf.write_text(text)
    sim = ref.Simulation(pf)
    f = sim.fields()
    sim.close()
    if half is None:
        # the box as the parameter file's unit conversion gives it (anchor = -half_pc pc, sides = 2 half_pc pc)
        a = ref.convert(-half_pc, "pc", "m")
        cs = ref.convert(2 * half_pc, "pc", "m") / nc
        m = a + cs * np.arange(nc) + 0.5 * cs
    else:
        cs = 2 * half / nc
        m = -half + cs * np.arange(nc) + 0.5 * cs      # CartesianDensityGrid::get_cell_midpoint
    X, Y, Z = np.meshgrid(m, m, m, indexing="ij")
    x = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 1)
    dens, temp, xH = host.ParameterFile(pf).density_function(x)
    assert np.array_equal(dens, f[0]) and np.array_equal(temp, f[1]) and np.array_equal(xH, f[2])
    assert np.isfinite(dens).all()
    if kind in ANALYTIC_PROFILES:
        assert np.unique(dens).size > 5 and dens.max() > 0
    if kind == "bondi_ionised_core":   # cells inside and outside the ionised core: two temperatures
        assert 20 < (temp > 4000.).sum() < temp.size and abs(temp.max() / temp.min() - 16.) < 1e-6   # P contrast 32, mu 1/2
    if kind in ("ascii_file", "interpolated_1d"):
        assert np.unique(dens).size > 20 and dens.min() >= 1e6 and dens.max() <= 3e8
    if kind == "lexington_blocks":
        assert (dens == 0).sum() > 0 and (dens > 0).sum() > 0


def test_uniform_and_faucher_giguere_spectra_from_a_parameter_file(host, ref, tmp_path, monkeypatch):
    """PhotonSourceSpectrumFactory types Uniform and FaucherGiguere in the host layer: the table the
    host builds from the fg_uvb_dec11 data files (frequencies, cumulative distribution, total flux)
    is the reference object's, bit for bit, at redshifts on the tables' 0.05 grid (in between the
    reference reads its second table from the wrong stream, host/IonizationSimulation.hpp)."""
    from pathlib import Path
    data = Path(__file__).resolve().parents[1] / "oracle" / "_ref" / "data"
    monkeypatch.setenv("CMIB_DATA_DIR", str(data))
    for z in (0., 0.45, 3.35, 7., 10.65):
        pf = tmp_path / f"fg_{z}.param"
        pf.write_text(f"ContinuousPhotonSourceSpectrum:\n  type: FaucherGiguere\n  redshift: {z}\n")
        p = host.ParameterFile(pf)
        s = p.photon_source_spectrum("ContinuousPhotonSourceSpectrum")
        p.close()
        d = ref.faucher_giguere(z, 1)
        assert s["kind"] == 3 and s["freq"].size == 100
        assert np.array_equal(s["freq"], d["freq"])
        assert np.array_equal(s["cdf"], d["cdf"]), z
        assert s["total_flux"] == d["total_flux"]
    # between two tables: a proper interpolation, bracketed by its neighbours
    fluxes = []
    for z in (3.35, 3.37, 3.4):
        pf = tmp_path / f"fg_{z}.param"
        pf.write_text(f"PhotonSourceSpectrum:\n  type: FaucherGiguere\n  redshift: {z}\n")
        p = host.ParameterFile(pf)
        fluxes.append(p.photon_source_spectrum()["total_flux"])
        p.close()
    assert min(fluxes[0], fluxes[2]) < fluxes[1] < max(fluxes[0], fluxes[2])
    assert abs(fluxes[1] - (0.6 * fluxes[0] + 0.4 * fluxes[2])) < 1e-9 * fluxes[1]
    pf = tmp_path / "uniform.param"
    pf.write_text("PhotonSourceSpectrum:\n  type: Uniform\n")
    p = host.ParameterFile(pf)
    assert p.photon_source_spectrum()["kind"] == 2
    p.close()
    pf = tmp_path / "bad.param"
    pf.write_text("PhotonSourceSpectrum:\n  type: WMBasic\n")
    p = host.ParameterFile(pf)
    with pytest.raises(Exception, match="Unknown PhotonSourceSpectrum type"):
        p.photon_source_spectrum()
    p.close()


def test_random_generator_stream_is_the_reference_stream(host, ref):
    """host/RandomGenerator.hpp (RANLUX level 2) against the compiled reference generator: every
    deviate, several seeds including the special cases 0 (-> 1) and negative seeds."""
    for seed in (42, 1, 0, 123456789, -7, 2 ** 31 - 1):
        assert np.array_equal(host.random_stream(seed, 100000), ref.random_stream(seed, 100000)), seed


def test_photon_source_distributions_give_the_reference_sources(host, ref, tmp_path):
    """PhotonSourceDistribution types of the host layer against the reference's factory on the same
    parameter file: positions, weights, total luminosity, bit for bit (UniformRandom draws its
    positions from the RANLUX stream and is evolved to the starting time)."""
    yml = tmp_path / "sources.yml"
    yml.write_text("number of sources: 3\nsource[0]:\n  position: [0. pc, 1. pc, -2. pc]\n  luminosity: 1.e49 s^-1\n"
                   "source[1]:\n  position: [1.e17 m, 0. m, 3. pc]\n  luminosity: 3.e48 s^-1\n"
                   "source[2]:\n  position: [-4. pc, -4. pc, 4. pc]\n  luminosity: 4.26e49 s^-1\n")
    cases = {
        "ascii": f"PhotonSourceDistribution:\n  type: AsciiFile\n  filename: {yml}\n",
        "uniform_defaults": "PhotonSourceDistribution:\n  type: UniformRandom\n  number of sources: 24\n",
        "uniform_evolved": ("PhotonSourceDistribution:\n  type: UniformRandom\n  number of sources: 50\n  random seed: 1234\n"
                            "  box anchor: [-3. pc, -2. pc, -1. pc]\n  box sides: [6. pc, 4. pc, 2. pc]\n"
                            "  source lifetime: 2. Myr\n  source luminosity: 3.e48 s^-1\n  update interval: 0.05 Myr\n"
                            "  starting time: 3.1 Myr\n"),
        # sources born at random and evolved to the starting time (rectangle x Gaussian height / Gaussian blob),
        # and the SILCC set (positions drawn when asked for): all on the RANLUX stream
        "disc_patch_defaults": "PhotonSourceDistribution:\n  type: DiscPatch\n",
        "disc_patch_evolved": ("PhotonSourceDistribution:\n  type: DiscPatch\n  source lifetime: 5. Myr\n"
                               "  source luminosity: 1.e49 s^-1\n  average number of sources: 40\n  anchor x: -200. pc\n"
                               "  sides x: 400. pc\n  anchor y: -100. pc\n  sides y: 300. pc\n  origin z: 10. pc\n"
                               "  scaleheight z: 40. pc\n  random seed: 77\n  update interval: 0.2 Myr\n"
                               "  starting time: 12.3 Myr\n"),
        "dwarf_galaxy_defaults": "PhotonSourceDistribution:\n  type: DwarfGalaxy\n",
        "dwarf_galaxy_evolved": ("PhotonSourceDistribution:\n  type: DwarfGalaxy\n  source lifetime: 8. Myr\n"
                                 "  average number of sources: 30\n  center: [10. pc, 20. pc, 30. pc]\n"
                                 "  scale radius: 150. pc\n  random seed: 5\n  update interval: 0.5 Myr\n"
                                 "  starting time: 0.031 Gyr\n"),
        "silcc_defaults": "PhotonSourceDistribution:\n  type: SILCC\n",
        "silcc": ("PhotonSourceDistribution:\n  type: SILCC\n  number of sources: 100\n  anchor x: -0.5 kpc\n"
                  "  sides x: 1. kpc\n  anchor y: -0.25 kpc\n  sides y: 0.5 kpc\n  origin z: 5. pc\n"
                  "  scaleheight z: 30. pc\n  luminosity: 1.e48 s^-1\n  random seed: 9\n"),
        "single": "PhotonSourceDistribution:\n  type: SingleStar\n  position: [1. pc, 2. pc, 3. pc]\n  luminosity: 2.e49 s^-1\n",
    }
    for name, text in cases.items():
        pf = tmp_path / f"{name}.param"
        pf.write_text(text)
        rp, rw, rl = ref.photon_source_distribution(pf)
        p = host.ParameterFile(pf)
        hp, hw, hl = p.photon_source_distribution()
        p.close()
        assert hp.shape == rp.shape and len(hp) > 0, name
        assert np.array_equal(hp, rp) and np.array_equal(hw, rw) and hl == rl, name
    assert len(hp) == 1


def test_fractal_density_mask_gives_the_reference_grid(host, ref, tmp_path):
    """DensityMask: Fractal — the masked initial grid of the host layer against the grid of the
    reference's IonizationSimulation::initialize on the same file, cell for cell, bit for bit
    (mask box smaller than the simulation box, partly smooth gas)."""
    nc = 16
    text = ("SimulationBox:\n  anchor: [-5. pc, -5. pc, -5. pc]\n  sides: [10. pc, 10. pc, 10. pc]\n"
            "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n"
            f"  number of cells: [{nc}, {nc}, {nc}]\nDensityFunction:\n  type: Homogeneous\n  density: 50. cm^-3\n"
            "DensityMask:\n  type: Fractal\n  box anchor: [-4. pc, -5. pc, -3. pc]\n  box sides: [8. pc, 10. pc, 7. pc]\n"
            "  resolution: [12, 10, 8]\n  number of particles: 20000\n  random seed: 77\n  fractal dimension: 2.4\n"
            "  number of levels: 3\n  fractal fraction: 0.8\nPhotonSourceSpectrum:\n  type: Monochromatic\n")
    pf = tmp_path / "fractal.param"
    pf.write_text(text)
    sim = ref.Simulation(pf)
    f = sim.fields()
    sim.close()
    p = host.ParameterFile(pf)
    dens = p.initial_number_density(nc ** 3)
    p.close()
    assert np.array_equal(dens, f[0])
    assert np.unique(dens).size > 100 and abs(dens.sum() / (5e7 * nc ** 3) - 1.) < 1e-12   # clumpy, atoms conserved


@pytest.mark.parametrize("mapping,periodic", [("centroid", False), ("centroid", True), ("M_over_V", True), ("M_over_V", False)])
def test_sph_array_interface_maps_like_the_reference(host, ref, tmp_path, mapping, periodic):
    """The device-free half of the coarse C ABI (cmi_compute_neutral_fraction_*): SPH particles ->
    densities on the parameter file's Cartesian grid (SPHArrayInterface::operator()) and a
    neutral-fraction field -> particles (the inverse mapping of SPHArrayInterface::write), against
    the reference's SPHArrayInterface on the same arrays.  The reference walks an octree, the host
    layer scatters particles over cells: same pairs, other summation order."""
    nc = 12
    pf = tmp_path / "sph.param"
    pf.write_text("SimulationBox:\n  anchor: [-5. pc, -5. pc, -5. pc]\n  sides: [10. pc, 10. pc, 10. pc]\n"
                  f"  periodicity: [{'true' if periodic else 'false'}, {'true' if periodic else 'false'}, "
                  f"{'true' if periodic else 'false'}]\nDensityGrid:\n  type: Cartesian\n"
                  f"  number of cells: [{nc}, {nc}, {nc}]\nPhotonSourceSpectrum:\n  type: Monochromatic\n")
    rng = np.random.default_rng(12)
    N = 3000
    pos = rng.uniform(-5 * PC, 5 * PC, (N, 3))
    pos[:50] = rng.uniform(-4.9 * PC, -4.0 * PC, (50, 3))       # a clump in a corner (wraps when periodic)
    h = np.exp(rng.uniform(np.log(0.3 * PC), np.log(1.6 * PC), N))
    m = np.full(N, 4.9e30) if mapping == "M_over_V" else rng.uniform(1e30, 9e30, N)
    xH = np.exp(rng.uniform(np.log(1e-5), 0., nc ** 3))
    box = ([-5 * PC] * 3, [10 * PC] * 3) if periodic else None
    args = (pf, mapping, pos[:, 0], pos[:, 1], pos[:, 2], h, m, nc ** 3, xH)
    rd, rn = ref.sph_mapping(*args, box=box)
    hd, hn = host.sph_mapping(*args, box=box)
    assert np.abs(hd / rd - 1.).max() < 1e-13
    assert np.abs(hn - rn).max() < 1e-12
    if mapping == "centroid":
        assert np.unique(hd).size > 0.9 * nc ** 3 and (hn < 0.999).mean() > 0.5   # small-h particles reach no midpoint
        # mean density ~ total mass / volume (SPH estimate at 1728 sample points)
        assert abs(hd.mean() * 1.6737236e-27 * (10 * PC) ** 3 / m.sum() - 1.) < 0.1
    else:
        assert np.unique(hd).size == 1 and np.array_equal(hn, rn)


def test_solar_metallicity_abundances_and_bimodal_cross_sections(host, ref, tmp_path):
    """AbundanceModel: SolarMetallicity and CrossSections: Bimodal from a parameter file, against the
    reference's factories: the six abundances and sigma[14](nu) on both sides of the frequency limit,
    bit for bit — including the reference's swapped oxygen_0 / sulphur_1 members and its odd
    "frequency limit:" key."""
    for Z in (None, -3.31, -3.0, -4.5):
        pf = tmp_path / f"abund_{Z}.param"
        pf.write_text("AbundanceModel:\n  type: SolarMetallicity\n" + ("" if Z is None else f"  metallicity: {Z}\n"))
        p = host.ParameterFile(pf)
        a = p.abundances()
        p.close()
        assert np.array_equal(a, ref.abundances(pf)), Z
        assert 0.08 < a[0] < 0.09 and (a[1:] > 0).all()
    rng = np.random.default_rng(2)
    keys = ["hydrogen_0", "helium_0", "carbon_1", "carbon_2", "nitrogen_0", "nitrogen_1", "nitrogen_2", "oxygen_0",
            "oxygen_1", "neon_0", "neon_1", "sulphur_1", "sulphur_2", "sulphur_3"]
    body = "".join(f"  {k}_low: {float(rng.uniform(1, 9)):.6f}e-18 cm^2\n  {k}_high: {float(rng.uniform(1, 9)):.6f}e-19 cm^2\n"
                   for k in keys)
    pf = tmp_path / "bimodal.param"
    pf.write_text("frequency limit: 20. eV\nCrossSections:\n  type: Bimodal\n" + body)
    nu = 3.288e15 * np.array([1.0, 1.05, 1.1029, 1.1030, 1.2, 1.4705, 1.4706, 1.5, 2.0, 3.9])
    p = host.ParameterFile(pf)
    sig = p.cross_sections(nu)
    p.close()
    r = ref.parameter_cross_sections(pf, nu)
    assert np.array_equal(sig, r)
    assert np.unique(sig[:, 0]).size == 2                       # both branches are exercised
    assert sig[0, 7] < sig[-1, 7] and sig[0, 0] > sig[-1, 0]    # oxygen_0 is swapped, hydrogen_0 is not


def test_planck_sampler_and_masked_spectrum_bitexact(host, ref, tmp_path):
    """(1) The Planck sampler the device runs (csrc/source.cuh planck_frequency_at), driven on the host by
    the RANLUX stream, returns the reference's frequencies bit for bit (PlanckPhotonSourceSpectrum.cpp:149-165).
    (2) PhotonSourceSpectrum: Masked (a Planck spectrum behind the Linear mask): the tabulated spectrum the
    host layer hands to the device — bins, cumulative distribution, total flux — is the reference's."""
    for T in (20000., 40000.):
        pf = tmp_path / f"planck_{T}.param"
        pf.write_text(f"PhotonSourceSpectrum:\n  type: Planck\n  temperature: {T} K\n")
        p = host.ParameterFile(pf)
        nu = p.sample_spectrum(200000, seed=17)
        p.close()
        assert np.array_equal(nu, ref.sample_spectrum(0, T, 200000, seed=17))
    pf = tmp_path / "masked.param"
    pf.write_text("PhotonSourceSpectrum:\n  type: Masked\n  masked type: Planck\n  temperature: 35000. K\n"
                  "  ionizing flux: 1.e12 m^-2 s^-1\n  mask number of bins: 500\n  mask number of samples: 2000000\n")
    p = host.ParameterFile(pf)
    s = p.photon_source_spectrum()
    p.close()
    r = ref.masked_spectrum(pf)
    assert s["kind"] == 3 and s["freq"].size == 500
    assert np.array_equal(s["freq"], r["freq"]) and np.array_equal(s["cdf"], r["cdf"])
    assert s["total_flux"] == r["total_flux"] and 0. < s["total_flux"] < 1e12
    assert s["cdf"][-1] == 1. and (np.diff(s["cdf"]) >= 0).all()


def test_multi_gpu_driver_threads_agree_on_a_failure(host):
    """The device threads of a multi-GPU iteration meet in a rendezvous before they enter the collective
    (IonizationSimulation.hpp): when one of them failed, ALL of them must learn it and skip the NCCL call (a rank that
    entered it alone would wait forever), and the next iteration starts clean."""
    for nthreads in (2, 3, 8):
        assert (host.test_rendezvous(nthreads, 4) == 1).all()
        for bad in (0, nthreads - 1):
            out = host.test_rendezvous(nthreads, 5, failing_thread=bad, failing_round=2)
            assert out.tolist() == [1, 1, 0, 1, 1], (nthreads, bad, out)


def test_malformed_parameter_files_end_in_errors_not_crashes(host, tmp_path):
    """A parameter file with broken indentation or stray characters is reported (the reference aborts with
    "Line has a different indentation than expected" / "no ':' found"); it must never take the process down:
    the reference's lexingtonHII20 file with 1-11 random edits, 200 mutants, in a child process."""
    import subprocess, sys
    from conftest import ROOT
    script = tmp_path / "fuzz.py"
    script.write_text(f"""
import sys, pathlib, numpy as np
sys.path.insert(0, {str(ROOT)!r})
from cmacionize_b200 import host
src = pathlib.Path({str(ROOT / 'tests' / 'golden' / 'benchmarks' / 'lexingtonHII20.param')!r}).read_text()
rng, tmp, errors = np.random.default_rng(5), pathlib.Path({str(tmp_path)!r}), 0
for k in range(200):
    b = list(src)
    for _ in range(int(rng.integers(1, 12))):
        p, m = int(rng.integers(0, len(b))), rng.integers(0, 4)
        if m == 0: b[p] = chr(int(rng.integers(32, 127)))
        elif m == 1: del b[p]
        elif m == 2: b.insert(p, "\\n")
        else: b.insert(p, ":")
    f = tmp / "mutant.param"
    f.write_text("".join(b))
    try:
        q = host.ParameterFile(f)
        try:
            q.photon_source_distribution(); q.abundances(); q.initial_number_density(64 ** 3)
        except (host.HostError, UnicodeDecodeError):
            errors += 1
        q.close()
    except (host.HostError, UnicodeDecodeError):
        errors += 1
print("survived", errors)
""")
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, (out.returncode, out.stderr[-1500:])
    assert out.stdout.strip().startswith("survived") and int(out.stdout.split()[-1]) > 20
    # the case that used to walk off the indentation stack
    bad = tmp_path / "bad.param"
    bad.write_text("A:\n    b: 1\n  c: 2\n")
    with pytest.raises(host.HostError, match="different indentation"):
        host.ParameterFile(bad)
