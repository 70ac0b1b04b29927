"""Gadget-style HDF5 snapshots written by the host layer (host/HDF5Writer.hpp, GadgetDensityGridWriter) without
an HDF5 library (reference: GadgetDensityGridWriter.cpp:122-300 through HDF5Tools.hpp / libhdf5).

There is no HDF5 library in the image either, so the chain of evidence is:
  1. tests/h5mini.py (a reader written from the file-format specification) is pinned on files written by the
     real library: the reference's test/test.hdf5 against the values its own testHDF5Tools.cpp asserts, and
     test/taskbased.hdf5, a snapshot written by the reference's GadgetDensityGridWriter;
  2. the writer's files are read back with that reader (every field bit for bit, every attribute);
  3. the writer's object-header messages are compared BYTE FOR BYTE with the messages the real library wrote for
     the same content (attributes of /Header, /Units, /RuntimePars in taskbased.hdf5; dataspace, datatype,
     fill-value and layout messages of a contiguous dataset in test/python_test.hdf5);
  4. a structural check of everything a reader follows (sizes, alignment, sorted symbol tables, B-tree keys,
     heap free list, end-of-file address).
"""
import struct
import subprocess
import sys

import numpy as np
import pytest

import h5mini
from conftest import ROOT

GOLD = ROOT / "tests" / "golden" / "hdf5"
PC = 3.086e16


@pytest.fixture(scope="module")
def host():
    subprocess.check_call([sys.executable, "-c", "from cmacionize_b200 import build as b; b.build_host()"], cwd=str(ROOT))
    from cmacionize_b200 import host as h
    return h


def test_reader_reads_what_the_reference_asserts_about_its_own_file():
    """test/testHDF5Tools.cpp:48-157 on test/test.hdf5"""
    f = h5mini.File(GOLD / "test.hdf5")
    assert "HydroScheme" in f and "NonExistingGroup" not in f
    g = f["HydroScheme"]
    assert len(g.attrs) == 15
    assert g.attrs["CFL parameter"][0] == np.float32(0.1) or abs(g.attrs["CFL parameter"][0] - 0.1) < 1e-7
    assert g.attrs["Dimension"][0] == 3
    assert g.attrs["Scheme"] == "Gadget-2 version of SPH (Springel 2005)"
    h = f["Header"]
    assert np.array_equal(h.attrs["BoxSize"], [1., 1., 1.])
    assert np.array_equal(h.attrs["NumPart_ThisFile"], [100, 0, 0, 0, 1, 0])
    assert np.array_equal(h.attrs["MassTable"], np.zeros(6))
    p = f["PartType0"]
    density = p["Density"].read()
    assert density.shape == (100,) and abs(density[0] / 0.12052436 - 1.) < 1e-8
    assert abs(density[31] - 0.2861183) < 1e-7 * 0.2861183 + 1e-9
    ids = p["ParticleIDs"].read()
    assert ids.shape == (100,) and ids[0] == 47
    x = p["Coordinates"].read()
    assert x.shape == (100, 3)
    assert np.allclose(x[0], [0.09859136052607954, 0.1422694476979986, 0.10086706479716455], rtol=1e-15, atol=0)


def test_reader_on_a_snapshot_written_by_the_reference():
    f = h5mini.File(GOLD / "taskbased.hdf5")
    assert f.link_order == ["Code", "Configuration", "Header", "Parameters", "PartType0", "RuntimePars", "Units"]
    h = f["Header"].attrs
    assert np.array_equal(h["BoxSize"], [3.086e17] * 3) and h["Dimension"] == 3 and h["NumFilesPerSnapshot"] == 1
    assert np.array_equal(h["NumPart_ThisFile"], [4096, 0, 0, 0, 0, 0])
    assert f["RuntimePars"].attrs["Iteration"] == 20
    assert f["Units"].attrs["Unit length in cgs (U_L)"] == 100.
    assert set(f["PartType0"].links()) == {"NeutralFractionH", "NumberDensity", "Temperature"}
    assert f["PartType0"]["NumberDensity"].space.shape == (4096,)
    assert f["Parameters"].attrs["SimulationBox:sides"] == "[3.086e+17 m, 3.086e+17 m, 3.086e+17 m]"


def _messages(obj, mtype):
    b = obj.f.buf
    return [bytes(b[body:body + size]) for t, body, size, _ in obj.messages if t == mtype]


def _attribute_message(obj, name):
    for m in _messages(obj, 0x000C):
        nsize = struct.unpack_from("<H", m, 2)[0]
        if m[8:8 + nsize].split(b"\0")[0].decode() == name:
            return m
    raise KeyError(name)


def check_structure(f):
    """Everything a reader follows, checked against the format rules (superblock v0 / object header v1)."""
    b = f.buf
    eof = len(b)
    assert f.eof == eof and f.base == 0 and f.leaf_k == 4 and f.internal_k == 16
    seen = []

    def check_header(o):
        version, _, nmsg, refcount, hsize = struct.unpack_from("<BBHII", b, o.addr)
        assert version == 1 and refcount == 1 and o.addr % 8 == 0
        assert o.addr + 16 + hsize <= eof
        total = 0
        for t, body, size, flags in o.messages:
            assert size % 8 == 0 and t != 0x0010        # one chunk, no continuation
            total += 8 + size
        assert total == hsize and len(o.messages) == nmsg
        seen.append((o.addr, o.addr + 16 + hsize))

    def check_group(o, path):
        check_header(o)
        btree, heap = o.symtab
        assert btree % 8 == 0 and heap % 8 == 0
        assert bytes(b[heap:heap + 4]) == b"HEAP" and b[heap + 4] == 0
        dsize, free, daddr = struct.unpack_from("<QQQ", b, heap + 8)
        assert daddr + dsize <= eof and dsize % 8 == 0
        assert bytes(b[daddr:daddr + 8]) == b"\0" * 8                 # the empty name of B-tree key 0
        assert free + 16 <= dsize
        nxt, fsize = struct.unpack_from("<QQ", b, daddr + free)
        assert nxt == 1 and free + fsize == dsize                     # one free block, up to the end
        assert bytes(b[btree:btree + 4]) == b"TREE"
        ntype, level, used = struct.unpack_from("<BBH", b, btree + 4)
        left, right = struct.unpack_from("<QQ", b, btree + 8)
        assert ntype == 0 and level == 0 and left == h5mini.UNDEF and right == h5mini.UNDEF and used <= 32
        seen.append((btree, btree + 24 + 33 * 8 + 32 * 8))
        seen.append((heap, heap + 32))
        seen.append((daddr, daddr + dsize))

        def name_at(off):
            e = b.index(b"\0", daddr + off)
            assert e < daddr + free
            return bytes(b[daddr + off:e])

        names = []
        prev_key = struct.unpack_from("<Q", b, btree + 24)[0]
        assert prev_key == 0
        for k in range(used):
            child, key = struct.unpack_from("<QQ", b, btree + 32 + 16 * k)
            assert bytes(b[child:child + 4]) == b"SNOD" and b[child + 4] == 1
            nsym = struct.unpack_from("<H", b, child + 6)[0]
            assert 1 <= nsym <= 8
            seen.append((child, child + 8 + 8 * 40))
            these = []
            for i in range(nsym):
                noff, oaddr, ctype, _ = struct.unpack_from("<QQII", b, child + 8 + 40 * i)
                these.append(name_at(noff))
                sub = h5mini.Obj(f, oaddr)
                if sub.symtab is not None:
                    assert ctype == 1 and struct.unpack_from("<QQ", b, child + 8 + 40 * i + 24) == sub.symtab
                else:
                    assert ctype == 0
            assert name_at(key) == these[-1]                          # key k+1 = largest name in child k
            if names:
                assert these[0] > names[-1]
            names += these
        assert names == sorted(names) and len(set(names)) == len(names)
        for nm, addr in o.links().items():
            sub = h5mini.Obj(f, addr)
            if sub.symtab is not None:
                check_group(sub, path + "/" + nm)
            else:
                check_header(sub)
                assert [t for t, *_ in sub.messages] == [0x0001, 0x0003, 0x0005, 0x0008, 0x0012]
                kind, a, s = sub.layout
                n = int(np.prod(sub.space.shape)) * sub.dtype.size
                assert kind == "contiguous" and s == n and a % 8 == 0 and a + n <= eof
                seen.append((a, a + ((n + 7) & ~7)))

    root_name_off, root_addr, ctype, _ = f.root_entry
    assert root_name_off == 0 and ctype == 1 and root_addr == 96
    assert struct.unpack_from("<QQ", b, 56 + 24) == f.symtab
    check_group(f, "")
    # the blocks tile the file exactly: no overlap, no hole
    seen.sort()
    assert seen[0][0] == 96
    for (a0, a1), (b0, b1) in zip(seen, seen[1:]):
        assert a1 == b0, (a0, a1, b0, b1)
    assert seen[-1][1] == eof


def _paramfile(tmp_path, ncell, extra=""):
    pf = tmp_path / "snap.param"
    pf.write_text("SimulationBox:\n  anchor: [-5. pc, -5. pc, -5. pc]\n  sides: [10. pc, 10. pc, 10. pc]\n"
                  "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n"
                  f"  number of cells: [{ncell[0]}, {ncell[1]}, {ncell[2]}]\n" + extra)
    return pf


def test_default_snapshot_is_the_reference_layout_and_reads_back_bit_for_bit(host, tmp_path):
    ncell = (6, 5, 4)
    n = int(np.prod(ncell))
    rng = np.random.default_rng(3)
    dens, T, x = rng.uniform(1e6, 1e9, n), rng.uniform(100., 3e4, n), rng.uniform(0., 1., (14, n))
    p = host.ParameterFile(_paramfile(tmp_path, ncell))
    name = p.write_snapshot(tmp_path, 20, dens, T, x)
    p.close()
    assert name == str(tmp_path / "snapshot020.hdf5")          # Utilities::compose_filename, padding 3
    f = h5mini.File(name)
    check_structure(f)
    ref = h5mini.File(GOLD / "taskbased.hdf5")
    assert f.link_order == ref.link_order                      # the reference's seven groups
    # default fields without hydro: Coordinates, NumberDensity, NeutralFractionH (DensityGridWriterFields.hpp:176-230)
    pt = f["PartType0"]
    assert set(pt.links()) == {"Coordinates", "NumberDensity", "NeutralFractionH"}
    assert np.array_equal(pt["NumberDensity"].read(), dens)
    assert np.array_equal(pt["NeutralFractionH"].read(), x[0])
    # Coordinates = cell midpoint - box anchor, cell order ix*ny*nz + iy*nz + iz
    cs = [10. * PC / k for k in ncell]
    ix, iy, iz = np.meshgrid(*[np.arange(k) for k in ncell], indexing="ij")
    mid = np.stack([(-5. * PC + cs[d] * i.reshape(-1) + 0.5 * cs[d]) - (-5. * PC) for d, i in enumerate((ix, iy, iz))], 1)
    assert np.array_equal(pt["Coordinates"].read(), mid)
    h = f["Header"].attrs
    assert list(f["Header"].attr_order) == list(ref["Header"].attr_order)
    assert np.array_equal(h["NumPart_ThisFile"], [n, 0, 0, 0, 0, 0]) and np.array_equal(h["NumPart_Total"], [n, 0, 0, 0, 0, 0])
    assert h["Time"] == 0.
    assert list(f["Units"].attr_order) == list(ref["Units"].attr_order) and f["Units"].attrs == ref["Units"].attrs
    assert list(f["RuntimePars"].attr_order) == ["Creation time", "Iteration"] and f["RuntimePars"].attrs["Iteration"] == 20
    par = f["Parameters"].attrs
    assert par["DensityGrid:number of cells"] == "[6, 5, 4]" and par["DensityGridWriter:type"] == "Gadget"
    assert par["SimulationBox:sides"] == ref["Parameters"].attrs["SimulationBox:sides"]
    # byte for byte what libhdf5 wrote for the same attributes (same box, same iteration)
    for group, names in (("Header", ["BoxSize", "Dimension", "Flag_Entropy_ICs", "MassTable", "NumFilesPerSnapshot",
                                     "NumPart_Total_HighWord", "Time"]),
                         ("Units", list(ref["Units"].attr_order)), ("RuntimePars", ["Iteration"]),
                         ("Parameters", ["SimulationBox:sides", "SimulationBox:periodicity"])):
        for nm in names:
            assert _attribute_message(f[group], nm) == _attribute_message(ref[group], nm), (group, nm)
    # a string attribute of the same length as the reference's time stamp has the same encoding around the text
    mine, theirs = _attribute_message(f["RuntimePars"], "Creation time"), _attribute_message(ref["RuntimePars"], "Creation time")
    assert len(mine) == len(theirs) and mine[:40] == theirs[:40]
    # group object headers: the symbol-table message has the library's encoding
    assert _messages(f["Header"], 0x0011)[0][:0] == b"" and len(_messages(f["Header"], 0x0011)[0]) == 16


def test_dataset_headers_are_byte_identical_to_a_contiguous_dataset_of_the_library(host, tmp_path):
    """test/python_test.hdf5 holds /PartType0/Coordinates (64 x 3 doubles, contiguous): a 4 x 4 x 4 grid gives the
    same shape, so dataspace, datatype and fill-value messages must be the same bytes, the layout message the same
    up to the address."""
    ncell = (4, 4, 4)
    p = host.ParameterFile(_paramfile(tmp_path, ncell))
    name = p.write_snapshot(tmp_path, 1, np.ones(64), np.ones(64), np.ones((14, 64)))
    p.close()
    f = h5mini.File(name)
    check_structure(f)
    mine, lib = f["PartType0"]["Coordinates"], h5mini.File(GOLD / "python_test.hdf5")["PartType0"]["Coordinates"]
    assert lib.layout[0] == "contiguous" and lib.space.shape == (64, 3)
    for mtype in (0x0001, 0x0003, 0x0005):
        assert _messages(mine, mtype) == _messages(lib, mtype), hex(mtype)
    a, b_ = _messages(mine, 0x0008)[0], _messages(lib, 0x0008)[0]
    assert a[:2] == b_[:2] and a[10:] == b_[10:] and len(a) == len(b_)   # version 3, contiguous, size 1536, padding
    flags = {t: fl for t, _, _, fl in mine.messages}
    assert flags == {t: fl for t, _, _, fl in lib.messages if t in flags}


def test_all_fields_and_the_ion_flag_quirk(host, tmp_path):
    ncell = (7, 3, 5)
    n = int(np.prod(ncell))
    rng = np.random.default_rng(8)
    dens, T, x = rng.uniform(1e6, 1e9, n), rng.uniform(100., 3e4, n), rng.uniform(0., 1., (14, n))
    # S+++ is the last ion: its flag switches on every ion before it (ion_present shifts the flag word)
    p = host.ParameterFile(_paramfile(tmp_path, ncell, "DensityGridWriter:\n  prefix: lex_\n  padding: 5\n"
                                      "DensityGridWriterFields:\n  Temperature: 1\n  NeutralFractionS+++: 1\n"))
    name = p.write_snapshot(tmp_path, 3, dens, T, x, time=2.5)
    p.close()
    assert name.endswith("lex_00003.hdf5")
    f = h5mini.File(name)
    check_structure(f)
    pt = f["PartType0"]
    ions = ["H", "He", "C+", "C++", "N", "N+", "N++", "O", "O+", "Ne", "Ne+", "S+", "S++", "S+++"]
    assert set(pt.links()) == {"Coordinates", "NumberDensity", "Temperature"} | {"NeutralFraction" + i for i in ions}
    assert len(f.btree_keys[pt.addr][0]) == 4                      # 17 links = 3 symbol nodes
    for k, ion in enumerate(ions):
        assert np.array_equal(pt["NeutralFraction" + ion].read(), x[k]), ion
    assert np.array_equal(pt["Temperature"].read(), T) and f["Header"].attrs["Time"] == 2.5
    # only He flagged: H (before it) comes along, nothing after it
    p = host.ParameterFile(_paramfile(tmp_path, ncell, "DensityGridWriterFields:\n  NeutralFractionH: 0\n  NeutralFractionHe: 1\n"
                                      "  Coordinates: 0\n"))
    name = p.write_snapshot(tmp_path, 0, dens, T, x)
    p.close()
    f = h5mini.File(name)
    check_structure(f)
    assert set(f["PartType0"].links()) == {"NumberDensity", "NeutralFractionH", "NeutralFractionHe"}


def test_writer_errors_and_ascii_type(host, tmp_path):
    ncell = (2, 2, 2)
    one = np.ones(8)
    p = host.ParameterFile(_paramfile(tmp_path, ncell, "DensityGridWriter:\n  type: AsciiFile\n"))
    name = p.write_snapshot(tmp_path, 2, one, one, np.ones((14, 8)))
    p.close()
    assert name.endswith("snapshot002.txt") and open(name).readline().startswith("#x (m)")
    for block, msg in (("DensityGridWriter:\n  type: Gadget\n  compression: true\n", "compression"),
                       ("DensityGridWriter:\n  type: Nope\n", "Unknown DensityGridWriter type"),
                       ("DensityGridWriterFields:\n  CosmicRayFactor: 1\n", "CosmicRayFactor")):
        p = host.ParameterFile(_paramfile(tmp_path, ncell, block))
        with pytest.raises(Exception, match=msg):
            p.write_snapshot(tmp_path, 0, one, one, np.ones((14, 8)))
        p.close()


# ---------------------------------------------------------------- reading snapshots (host/HDF5Reader.hpp)
def test_cpp_reader_equals_the_python_reader_on_files_of_the_real_library(host):
    """chunked + shuffle + deflate datasets of a snapshot written by the reference, chunked and attribute data of
    test/test.hdf5 (float32 / uint64 / 2-D), contiguous data of test/python_test.hdf5"""
    for name, paths in (("taskbased.hdf5", ["/PartType0/NumberDensity", "/PartType0/NeutralFractionH", "/PartType0/Temperature"]),
                        ("test.hdf5", ["/PartType0/Density", "/PartType0/ParticleIDs", "/PartType0/Coordinates",
                                       "/PartType0/Velocities", "/PartType4/Coordinates"]),
                        ("python_test.hdf5", ["/PartType0/Coordinates"])):
        f, h = h5mini.File(GOLD / name), host.HDF5Input(GOLD / name)
        for p in paths:
            a = f[p].read()
            b = h.dataset(p)
            assert a.shape == b.shape and np.array_equal(a.astype(np.float64), b), (name, p)
    h, f = host.HDF5Input(GOLD / "test.hdf5"), h5mini.File(GOLD / "test.hdf5")
    assert h.exists("/HydroScheme") and h.exists("/PartType0/Density") and not h.exists("/NonExistingGroup")
    assert h.attribute_names("/HydroScheme") == f["HydroScheme"].attr_order and len(h.attribute_names("/HydroScheme")) == 15
    assert h.string_attribute("/HydroScheme", "Scheme") == "Gadget-2 version of SPH (Springel 2005)"
    assert np.array_equal(h.numeric_attribute("/Header", "NumPart_ThisFile"), [100, 0, 0, 0, 1, 0])
    assert abs(h.numeric_attribute("/HydroScheme", "CFL parameter")[0] - 0.1) < 1e-7
    t = host.HDF5Input(GOLD / "taskbased.hdf5")
    assert t.string_attribute("/Parameters", "SimulationBox:sides") == "[3.086e+17 m, 3.086e+17 m, 3.086e+17 m]"
    x = t.dataset("/PartType0/NeutralFractionH")
    assert t.dataset("/PartType0/NumberDensity").tolist() == [1e8] * 4096 and 0. < x.min() < 1e-5 and x.max() == 1.
    with pytest.raises(Exception, match="does not exist"):
        h.dataset("/PartType0/Nope")
    with pytest.raises(Exception, match="not an HDF5 file"):
        host.HDF5Input(ROOT / "tests" / "golden" / "reference_fixtures.npz").exists("/x")


def _midpoints(ncell, half):
    cs = [2 * half / k for k in ncell]
    ix, iy, iz = np.meshgrid(*[np.arange(k) for k in ncell], indexing="ij")
    return np.stack([-half + cs[d] * i.reshape(-1) + 0.5 * cs[d] for d, i in enumerate((ix, iy, iz))], 1)


def test_snapshot_density_function_round_trip_and_resampling(host, tmp_path):
    """DensityFunction type CMacIonizeSnapshot on a snapshot of this host layer: the same grid gets its cells back
    bit for bit; a coarser grid gets the snapshot cell that contains each midpoint
    (CMacIonizeSnapshotDensityFunction.cpp:504-523)."""
    ncell = (6, 4, 8)
    n = int(np.prod(ncell))
    rng = np.random.default_rng(21)
    dens, T, x = rng.uniform(1e6, 1e9, n), rng.uniform(100., 3e4, n), rng.uniform(0., 1., (14, n))
    p = host.ParameterFile(_paramfile(tmp_path, ncell, "DensityGridWriterFields:\n  Temperature: 1\n  NeutralFractionHe: 1\n"))
    snap = p.write_snapshot(tmp_path, 4, dens, T, x)
    p.close()
    again = tmp_path / "again.param"
    again.write_text(_paramfile(tmp_path, ncell).read_text() + f"DensityFunction:\n  type: CMacIonizeSnapshot\n  filename: {snap}\n")
    p = host.ParameterFile(again)
    d2, T2, x2 = p.density_function(_midpoints(ncell, 5 * PC))
    p.close()
    assert np.array_equal(d2, dens) and np.array_equal(T2, T) and np.array_equal(x2, x[0])
    coarse = (3, 2, 2)
    p = host.ParameterFile(again)
    d3, T3, x3 = p.density_function(_midpoints(coarse, 5 * PC))
    p.close()
    m = _midpoints(coarse, 5 * PC)
    idx = [np.floor((m[:, d] + 5 * PC) / (10 * PC) * ncell[d]).astype(int) for d in range(3)]
    flat = (idx[0] * ncell[1] + idx[1]) * ncell[2] + idx[2]
    assert np.array_equal(d3, dens[flat]) and np.array_equal(T3, T[flat]) and np.array_equal(x3, x[0][flat])
    # a snapshot without temperatures cannot restart a run: said so
    p = host.ParameterFile(_paramfile(tmp_path, ncell))
    bare = p.write_snapshot(tmp_path, 9, dens, T, x)
    p.close()
    bad = tmp_path / "bad.param"
    bad.write_text(_paramfile(tmp_path, ncell).read_text() + f"DensityFunction:\n  type: CMacIonizeSnapshot\n  filename: {bare}\n")
    p = host.ParameterFile(bad)
    with pytest.raises(Exception, match="holds no Temperature"):
        p.density_function(_midpoints(ncell, 5 * PC))
    p.close()


def test_snapshot_density_function_on_a_snapshot_of_the_reference(host, tmp_path):
    """test/taskbased.hdf5: written by the reference's task-based driver (subgrid cell order, chunked + shuffled +
    deflated datasets): a Stromgren sphere comes back, ionised at the source and neutral in the corners."""
    f = h5mini.File(GOLD / "taskbased.hdf5")
    par = f["Parameters"].attrs
    nsub = [int(v) for v in par["DensitySubGridCreator:number of subgrids"].strip("[]").split(",")]
    ncell = [int(v) for v in par["DensityGrid:number of cells"].strip("[]").split(",")]
    assert "DensityGrid:type" not in par and ncell == [16, 16, 16]
    pf = tmp_path / "restart.param"
    pf.write_text(f"SimulationBox:\n  anchor: {par['SimulationBox:anchor']}\n  sides: {par['SimulationBox:sides']}\n"
                  "  periodicity: [false, false, false]\nDensityGrid:\n  type: Cartesian\n  number of cells: [16, 16, 16]\n"
                  f"DensityFunction:\n  type: CMacIonizeSnapshot\n  filename: {GOLD / 'taskbased.hdf5'}\n")
    half = 0.5 * 3.086e17
    m = _midpoints(ncell, half)
    p = host.ParameterFile(pf)
    dens, T, xH = p.density_function(m)
    p.close()
    assert (dens == 1e8).all() and (T == 8000.).all()
    raw = f["PartType0"]["NeutralFractionH"].read()
    # the subgrid order of TaskBased snapshots (CMacIonizeSnapshotDensityFunction.cpp:372-418)
    nb = [ncell[d] // nsub[d] for d in range(3)]
    grid = np.empty(ncell)
    k = 0
    for six in range(nsub[0]):
        for siy in range(nsub[1]):
            for siz in range(nsub[2]):
                blk = raw[k:k + nb[0] * nb[1] * nb[2]].reshape(nb)
                grid[six * nb[0]:(six + 1) * nb[0], siy * nb[1]:(siy + 1) * nb[1], siz * nb[2]:(siz + 1) * nb[2]] = blk
                k += blk.size
    assert np.array_equal(xH, grid.reshape(-1))
    r = np.sqrt((m ** 2).sum(1))
    assert xH[r < 0.3 * half].max() < 1e-3 and xH[r > 1.2 * half].min() > 0.9


# ---------------------------------------------------------------- SPH snapshots (DensityFunction GadgetSnapshot)
def _gadget_param(tmp_path, snapshot, anchor, sides, ncell, extra=""):
    pf = tmp_path / "gadget.param"
    anchor, sides = [float(v) for v in anchor], [float(v) for v in sides]
    pf.write_text(f"SimulationBox:\n  anchor: [{anchor[0]!r} m, {anchor[1]!r} m, {anchor[2]!r} m]\n"
                  f"  sides: [{sides[0]!r} m, {sides[1]!r} m, {sides[2]!r} m]\n  periodicity: [false, false, false]\n"
                  f"DensityGrid:\n  type: Cartesian\n  number of cells: [{ncell[0]}, {ncell[1]}, {ncell[2]}]\n"
                  f"DensityFunction:\n  type: GadgetSnapshot\n  filename: {snapshot}\n" + extra)
    return pf


def _grid_midpoints(anchor, sides, ncell):
    ax = [anchor[d] + sides[d] / ncell[d] * (np.arange(ncell[d]) + 0.5) for d in range(3)]
    X, Y, Z = np.meshgrid(*ax, indexing="ij")
    return np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], 1)


def test_gadget_snapshot_density_function_on_the_reference_test_file(host, ref, tmp_path):
    """test/test.hdf5 is the file of the reference's testGadgetSnapshotDensityFunction.cpp (100 gas particles,
    periodic unit box, SI-like units, float32 data, a Temperature dataset that was never written): the kernel sums
    at the cell midpoints against the reference's Octree + CubicSplineKernel composed as operator() does, and the
    assertions of that test (hydrogen number of the 32^3 grid = that of the particles; average temperature 0)."""
    f = h5mini.File(GOLD / "test.hdf5")
    gas = f["PartType0"]
    UL, UM = 100. * 0.01, 1000. * 0.001                           # Units group -> SI (cm, g)
    pos = gas["Coordinates"].read().astype(np.float64) * UL
    m = gas["Masses"].read().astype(np.float64) * UM
    h = gas["SmoothingLength"].read().astype(np.float64) * UL
    rho = gas["Density"].read().astype(np.float64) * (UM / UL / (UL * UL))
    T = gas["Temperature"].read().astype(np.float64)
    assert len(m) == 100 and (T == 0).all() and f["RuntimePars"].attrs["PeriodicBoundariesOn"][0] == 1
    ncell = (12, 12, 12)
    q = _grid_midpoints((0., 0., 0.), (1., 1., 1.), ncell)
    p = host.ParameterFile(_gadget_param(tmp_path, GOLD / "test.hdf5", (0., 0., 0.), (1., 1., 1.), ncell))
    dens, temp, xH = p.density_function(q)
    p.close()
    rd, rT, rx = ref.gadget_kernel_sums(pos, m, h, rho, np.ones(100), None, True, (1., 1., 1.), q)
    assert (rd > 0).all() and (rx == -1).all()
    assert np.abs(dens / rd - 1.).max() < 1e-13                    # same particles, another order of the sum
    assert (temp == 0.).all() and (xH == 1e-6).all()
    # testGadgetSnapshotDensityFunction.cpp:53-57 on its 32^3 grid
    ncell = (32, 32, 32)
    p = host.ParameterFile(_gadget_param(tmp_path, GOLD / "test.hdf5", (0., 0., 0.), (1., 1., 1.), ncell))
    dens, temp, _ = p.density_function(_grid_midpoints((0., 0., 0.), (1., 1., 1.), ncell))
    p.close()
    total = (dens * (1. / 32) ** 3).sum()
    assert abs(total / (m.sum() / 1.6737236e-27) - 1.) < 1e-4 and temp.mean() == 0.


@pytest.mark.parametrize("periodic", [False, True])
def test_gadget_snapshot_density_function_on_synthetic_particles(host, ref, tmp_path, periodic):
    """clustered particles with a wide range of smoothing lengths, temperatures and neutral fractions, unit
    conversion through /Units (kpc, 1e10 Msol: GADGET units), cells partly outside the particle cloud"""
    rng = np.random.default_rng(17)
    N = 3000
    KPC, M10 = 3.086e19, 1.98855e40
    box = np.array([8., 6., 5.])                                   # snapshot units (kpc)
    centres = rng.uniform(1., 4., (6, 3))
    pos = (centres[rng.integers(0, 6, N)] + rng.normal(0., 0.5, (N, 3)))
    if periodic:
        pos = np.mod(pos, box)
    h = rng.uniform(0.05, 0.9, N)
    m = rng.uniform(0.5, 2., N) * 1e-6
    rho = rng.uniform(0.1, 10., N) * 1e-4
    T = rng.uniform(1e2, 1e5, N)
    xH = rng.uniform(0., 1., N)
    snap = tmp_path / "sph.hdf5"
    host.write_particle_snapshot(snap, pos, m, h, rho, T, xH, periodic=1 if periodic else -1, boxsize=box,
                                 units_cgs=(KPC * 100., M10 * 1000., 1.))
    UL, UM = (KPC * 100.) * 0.01, (M10 * 1000.) * 0.001           # what UnitConverter::to_SI does with cm, g
    sp, sm, sh, srho = pos * UL, m * UM, h * UL, rho * (UM / UL / (UL * UL))
    anchor = (0., 0., 0.) if periodic else tuple((pos.min(0) - 0.3) * UL)
    sides = tuple(box * UL) if periodic else tuple((pos.max(0) - pos.min(0) + 0.6) * UL)
    ncell = (14, 11, 9)
    q = _grid_midpoints(anchor, sides, ncell)
    for use_x in (False, True):
        p = host.ParameterFile(_gadget_param(tmp_path, snap, anchor, sides, ncell,
                                             "  use neutral fraction: true\n" if use_x else ""))
        dens, temp, x = p.density_function(q)
        p.close()
        rd, rT, rx = ref.gadget_kernel_sums(sp, sm, sh, srho, T, xH if use_x else None, periodic, tuple(box * UL), q)
        inside = rd > 0
        assert inside.sum() > 0.3 * len(q) and (periodic or (~inside).sum() > 0)
        assert np.array_equal(dens == 0, ~inside)
        assert np.abs(dens[inside] / rd[inside] - 1.).max() < 1e-12
        assert np.abs(temp[inside] / rT[inside] - 1.).max() < 1e-12 and (temp[~inside] == 0).all()
        if use_x:
            assert np.abs(x[inside] / rx[inside] - 1.).max() < 1e-12
        else:
            assert (x == 1e-6).all()
        # the grid fill of a run (particles scattered over the cells) gives the same cells as the point queries
        p = host.ParameterFile(_gadget_param(tmp_path, snap, anchor, sides, ncell,
                                             "  use neutral fraction: true\n" if use_x else ""))
        gd, gT, gx = p.initial_grid(len(q))
        p.close()
        assert np.array_equal(gd == 0, dens == 0) and np.abs(gd[inside] / dens[inside] - 1.).max() < 1e-12
        assert np.abs(gT[inside] / temp[inside] - 1.).max() < 1e-12 and (gT[~inside] == 0).all()
        assert np.abs(gx[inside] / x[inside] - 1.).max() < 1e-12


def test_gadget_snapshot_source_distribution(host, tmp_path):
    """testGadgetSnapshotPhotonSourceDistribution.cpp:89-107 on test/test.hdf5 (one star of unit mass in the middle
    of the unit box, RateBased luminosity function with rate 2 and cutoff age 2 -> one source of luminosity 2), then
    stars and star-forming gas of a synthetic snapshot against the rule written out in numpy."""
    pf = tmp_path / "stars.param"
    pf.write_text("SimulationBox:\n  anchor: [0. m, 0. m, 0. m]\n  sides: [1. m, 1. m, 1. m]\n"
                  f"PhotonSourceDistribution:\n  type: GadgetSnapshot\n  filename: {GOLD / 'test.hdf5'}\n"
                  "UVLuminosityFunction:\n  type: RateBased\n  UV rate per mass unit: 2. s^-1 kg^-1\n  cutoff age: 2. s\n")
    p = host.ParameterFile(pf)
    pos, w, lum = p.photon_source_distribution()
    p.close()
    assert pos.tolist() == [[0.5, 0.5, 0.5]] and w.tolist() == [1.] and lum == 2.
    rng = np.random.default_rng(5)
    N, NS = 400, 300
    KPC, M10, MYR = 3.086e19, 1.98855e40, 3.154e13
    gpos, spos = rng.uniform(-1., 11., (N, 3)), rng.uniform(-1., 11., (NS, 3))
    sfr = np.where(rng.uniform(size=N) < 0.3, rng.uniform(0.1, 2., N), 0.)
    form, smass = rng.uniform(0., 30., NS), rng.uniform(1e-7, 1e-6, NS)
    snap = tmp_path / "galaxy.hdf5"
    host.write_particle_snapshot(snap, gpos, np.ones(N), np.ones(N), np.ones(N), units_cgs=(KPC * 100., M10 * 1000., 1.),
                                 unit_time_cgs=MYR, time=30., sfr=sfr, stars=(spos, form, smass))
    UL, UM, UT = (KPC * 100.) * 0.01, (M10 * 1000.) * 0.001, MYR
    box_lo, box_hi = np.zeros(3), np.full(3, 10. * KPC)
    rate, cut = 2.49428e16, 5. * MYR
    for use_gas in (False, True):
        pf.write_text("SimulationBox:\n  anchor: [0. kpc, 0. kpc, 0. kpc]\n  sides: [10. kpc, 10. kpc, 10. kpc]\n"
                      f"PhotonSourceDistribution:\n  type: GadgetSnapshot\n  filename: {snap}\n"
                      + ("  use gas: true\n" if use_gas else ""))
        p = host.ParameterFile(pf)
        pos, w, lum = p.photon_source_distribution()
        p.close()
        if use_gas:
            x = gpos * UL
            L = np.where(sfr > 0, sfr * (UM / UT) * cut * rate, 0.)
        else:
            x = spos * UL
            L = np.where((30. - form) * UT <= cut, smass * UM * rate, 0.)
        keep = (L > 0) & ((x >= box_lo) & (x < box_hi)).all(1)
        assert 5 < keep.sum() < len(keep)
        assert np.array_equal(pos, x[keep])
        tot = 0.
        for v in L[keep]:
            tot += v
        assert lum == tot and np.array_equal(w, L[keep] / tot)


def test_flash_snapshot_density_function_reproduces_the_reference_test(host, tmp_path):
    """testFLASHSnapshotDensityFunction.cpp:52-73 on test/FLASHtest.hdf5 (82 blocks of 8^3 cells, float32 data, the box
    and the block counts in {name, value} compound tables): the same 128 positions, the same statistic against the
    analytic density the file was made from (the reference asserts xi2 = 0.0106294), temperature 4000 K everywhere.
    Then the grid fill of a run against the point queries, cell for cell."""
    f = h5mini.File(GOLD / "FLASHtest.hdf5")
    assert f["node type"].read().shape == (82,) and f["dens"].space.shape == (82, 8, 8, 8)
    real = {r["name"].decode().strip(): float(r["value"]) for r in f["real runtime parameters"].read()}
    integer = {r["name"].decode().strip(): int(r["value"]) for r in f["integer runtime parameters"].read()}
    assert (real["xmin"], real["xmax"], real["ymax"], real["zmax"]) == (0., 2., 1., 1.) and integer["nblockx"] == 2
    pf = tmp_path / "flash.param"
    # (a box slightly inside the snapshot's, so that no cell midpoint sits exactly on a FLASH cell boundary)
    pf.write_text("SimulationBox:\n  anchor: [1.3e-4 m, 0.7e-4 m, 0.9e-4 m]\n  sides: [0.0195 m, 0.0097 m, 0.0096 m]\n"
                  "  periodicity: [false, false, false]\n"
                  "DensityGrid:\n  type: Cartesian\n  number of cells: [21, 13, 11]\n"
                  f"DensityFunction:\n  type: FLASHSnapshot\n  filename: {GOLD / 'FLASHtest.hdf5'}\n")
    i = np.arange(128)
    q = np.stack([(i + 0.5) * 0.02 / 128, (i + 0.5) * 0.01 / 128, (i + 0.5) * 0.01 / 128], 1)
    p = host.ParameterFile(pf)
    dens, T, xH = p.density_function(q)
    expected = (1. + 100. * q[:, 0] + 100. * q[:, 1] + 100. * q[:, 2]) * 1.e3 / 1.6737236e-27
    xi2 = (((dens - expected) / (dens + expected)) ** 2).sum()
    assert abs(xi2 / 0.0106294 - 1.) < 1e-5 and (T == 4000.).all() and (xH == 1e-6).all()
    m = _grid_midpoints((1.3e-4, 0.7e-4, 0.9e-4), (0.0195, 0.0097, 0.0096), (21, 13, 11))
    d1, T1, x1 = p.density_function(m)
    d2, T2, x2 = p.initial_grid(len(m))
    p.close()
    assert np.array_equal(d1, d2) and np.array_equal(T1, T2) and np.array_equal(x1, x2) and np.unique(d2).size > 50
    pf.write_text(pf.read_text() + "  temperature: 7500. K\n")
    p = host.ParameterFile(pf)
    assert (p.initial_grid(len(m))[1] == 7500.).all()
    p.close()
    pf.write_text(pf.read_text().replace("sides: [0.0195 m,", "sides: [0.03 m,"))
    p = host.ParameterFile(pf)
    with pytest.raises(Exception, match="lies outside the blocks"):
        p.initial_grid(len(m))
    p.close()


@pytest.mark.parametrize("hrange", [(0.05, 0.3), (0.3, 0.45), (0.5, 0.9)])
def test_gadget_snapshot_periodic_box_with_kernels_as_large_as_the_box(host, ref, tmp_path, hrange):
    """periodic unit box, smoothing lengths up to 0.9 of the box side (3, 2 and 1 search bins per axis; beyond half the
    box only the nearest image counts, Box::periodic_distance): point queries and the grid fill against the reference's
    Octree + CubicSplineKernel"""
    rng = np.random.default_rng(3)
    N = 300
    pos, h = rng.uniform(0., 1., (N, 3)), rng.uniform(hrange[0], hrange[1], N)
    m, rho, T, xH = rng.uniform(1., 2., N), rng.uniform(1., 2., N), rng.uniform(10., 100., N), rng.uniform(0., 1., N)
    snap = tmp_path / "periodic.hdf5"
    host.write_particle_snapshot(snap, pos, m, h, rho, T, xH, periodic=1, boxsize=(1., 1., 1.))
    ncell = (7, 6, 5)
    q = _grid_midpoints((0., 0., 0.), (1., 1., 1.), ncell)
    p = host.ParameterFile(_gadget_param(tmp_path, snap, (0., 0., 0.), (1., 1., 1.), ncell, "  use neutral fraction: true\n"))
    dens, temp, x = p.density_function(q)
    gd, gT, gx = p.initial_grid(len(q))
    p.close()
    rd, rT, rx = ref.gadget_kernel_sums(pos, m, h, rho, T, xH, True, (1., 1., 1.), q)
    for mine, theirs in ((dens, rd), (temp, rT), (x, rx), (gd, rd), (gT, rT), (gx, rx)):
        assert np.abs(mine / theirs - 1.).max() < 1e-13


FUZZ = r"""
import sys, pathlib, numpy as np
sys.path.insert(0, sys.argv[1])
from cmacionize_b200 import host
src, tmp, rng = pathlib.Path(sys.argv[2]).read_bytes(), pathlib.Path(sys.argv[3]), np.random.default_rng(int(sys.argv[4]))
errors = 0
for k in range(int(sys.argv[5])):
    b = bytearray(src)
    mode = rng.integers(0, 3)
    if mode == 0:
        b = b[:int(rng.integers(96, len(b)))]
    elif mode == 1:
        for _ in range(int(rng.integers(1, 8))):
            b[int(rng.integers(8, min(len(b), 20000)))] = int(rng.integers(0, 256))
    else:
        for _ in range(int(rng.integers(1, 4))):
            p = int(rng.integers(8, min(len(b), 20000) - 8))
            b[p:p + 8] = rng.integers(0, 256, 8, dtype=np.uint8).tobytes()
    f = tmp / f"mutant{k}.hdf5"
    f.write_bytes(bytes(b))
    h = host.HDF5Input(f)
    for path in ("/PartType0/Coordinates", "/PartType0/Density", "/PartType0/NumberDensity", "/PartType0/Temperature", "/dens"):
        try:
            if h.exists(path):
                h.dataset(path)
        except (host.HostError, UnicodeDecodeError):
            errors += 1
    for g in ("/Header", "/Parameters", "/Units"):
        try:
            for a in (h.attribute_names(g)[:5] if h.exists(g) else []):
                try:
                    h.numeric_attribute(g, a)
                except (host.HostError, UnicodeDecodeError):
                    errors += 1
        except (host.HostError, UnicodeDecodeError):
            errors += 1
    f.unlink()
print("survived", errors)
"""


@pytest.mark.parametrize("name", ["test.hdf5", "taskbased.hdf5", "FLASHtest.hdf5"])
def test_reader_survives_truncated_and_corrupted_files(host, tmp_path, name):
    """Snapshots come from other codes and from interrupted runs: a truncated file, flipped bytes or garbage addresses
    must end in an error message, never in a crash of the process that holds the GPU context (every offset and size the
    reader follows is checked against the mapped file).  80 mutants per file, in a child process."""
    script = tmp_path / "fuzz.py"
    script.write_text(FUZZ)
    out = subprocess.run([sys.executable, str(script), str(ROOT), str(GOLD / name), str(tmp_path), "7", "80"],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, (out.returncode, out.stderr[-1500:])
    assert out.stdout.strip().startswith("survived") and int(out.stdout.split()[-1]) > 0


TARGETED = r"""
import sys, pathlib, struct
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, sys.argv[1] + "/tests")
import h5mini
from cmacionize_b200 import host
src, tmp = pathlib.Path(sys.argv[2]), pathlib.Path(sys.argv[3])
f = h5mini.File(src)
chunked = [n for n in f["PartType0"].links() if f["PartType0"][n].layout and f["PartType0"][n].layout[0] == "chunked"]
assert chunked, "the fixture has a chunked dataset"
errors = 0
for name in chunked[:2]:
    d = f["PartType0"][name]
    node = d.layout[1]
    raw = src.read_bytes()
    assert raw[node:node + 4] == b"TREE" and raw[node + 4] == 1
    nd = len(d.layout[2]) - 1
    key_size = 8 + 8 * (nd + 1)
    # (1) a node that claims to be an internal node and whose first child is the node itself: a cycle
    b = bytearray(raw); b[node + 5] = 1; b[node + 24 + key_size:node + 24 + key_size + 8] = struct.pack("<Q", node)
    # (2) a dataspace of 2^32 x 2^32 (x ...) elements: the element count wraps to 0
    space = [body for t, body, _, _ in d.messages if t == 0x0001][0]
    c = bytearray(raw); p0 = space + (8 if c[space] == 1 else 4)
    for k in range(c[space + 1]): c[p0 + 8 * k:p0 + 8 * k + 8] = struct.pack("<Q", 1 << 32)
    # (3) huge but not wrapping: must be an error message, not std::bad_alloc / an out-of-bounds scatter
    e = bytearray(raw); e[p0:p0 + 8] = struct.pack("<Q", 1 << 40)
    # (4) a chunk offset outside the dataspace
    g = bytearray(raw); g[node + 24 + 8:node + 24 + 16] = struct.pack("<Q", (1 << 63) + 5)
    for k, m in enumerate((b, c, e, g)):
        mf = tmp / f"targeted{k}.hdf5"
        mf.write_bytes(bytes(m))
        try:
            host.HDF5Input(mf).dataset("/PartType0/" + name)
        except host.HostError as err:
            errors += 1
        mf.unlink()
print("survived", errors)
"""


def test_reader_refuses_cyclic_chunk_trees_and_overflowing_dataspaces(host, tmp_path):
    """The two crashes the round-1 review found by hand (a chunk B-tree node whose child is itself: stack overflow; a
    dataspace of 2^32 x 2^32 elements whose product wraps to 0: out-of-bounds scatter), plus a huge dataspace and a chunk
    offset outside the dataspace: each must raise the reader's error in a process that keeps running."""
    script = tmp_path / "targeted.py"
    script.write_text(TARGETED)
    out = subprocess.run([sys.executable, str(script), str(ROOT), str(GOLD / "test.hdf5"), str(tmp_path)],
                         capture_output=True, text=True, timeout=120)
    assert out.returncode == 0, (out.returncode, out.stderr[-1500:])
    n = int(out.stdout.split()[-1])
    assert n >= 4 and n % 4 == 0, out.stdout   # every mutant of every chunked dataset ended in an error


def test_benchmark_parameter_file_writes_the_fields_its_analysis_script_reads(host, tmp_path):
    """benchmarks/lexingtonHII20.param (unmodified copy, grid shrunk): the snapshot is named as benchmarks/lexingtonHII20.py
    globs it and holds exactly the datasets that script opens (Coordinates, Temperature, NeutralFraction<ion> of the 14
    ions: its DensityGridWriterFields block switches NumberDensity off and the rest on)."""
    text = (ROOT / "tests" / "golden" / "benchmarks" / "lexingtonHII20.param").read_text()
    assert "[64, 64, 64]" in text
    pf = tmp_path / "lexingtonHII20.param"
    pf.write_text(text.replace("[64, 64, 64]", "[4, 4, 4]"))
    p = host.ParameterFile(pf)
    name = p.write_snapshot(tmp_path, 20, np.ones(64), np.full(64, 7500.), np.full((14, 64), 0.25))
    p.close()
    assert name.endswith("/lexingtonHII20_020.hdf5")
    f = h5mini.File(name)
    check_structure(f)
    ions = ["H", "He", "C+", "C++", "N", "N+", "N++", "O", "O+", "Ne", "Ne+", "S+", "S++", "S+++"]
    assert set(f["PartType0"].links()) == {"Coordinates", "Temperature"} | {"NeutralFraction" + i for i in ions}
    assert (f["PartType0"]["Temperature"].read() == 7500.).all() and np.array_equal(f["Header"].attrs["BoxSize"], [6 * PC] * 3)
