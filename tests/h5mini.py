"""Minimal HDF5 reader (test infrastructure: there is no HDF5 library in this image).

Reads what the reference's snapshot files and `host/HDF5Writer.hpp` contain: superblock version 0,
symbol-table groups (B-tree v1 + SNOD + local heap), version-1 object headers with continuation
blocks, attributes (version 1 messages), fixed-point / IEEE float / fixed-length string datatypes and compounds of
those,
simple and scalar dataspaces, contiguous, compact and unfiltered chunked layouts.

Pinned on files written by the real library: `tests/golden/hdf5/` holds copies of the reference's own
`test/test.hdf5` (the values its `testHDF5Tools.cpp` asserts) and `test/taskbased.hdf5` (a CMacIonize
snapshot written by the reference's GadgetDensityGridWriter); `tests/test_hdf5_writer.py` checks the
reader on those before it trusts it on the files of the host layer's writer.
"""
from __future__ import annotations

import struct

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(RuntimeError):
    pass


def _pad8(n):
    return (n + 7) & ~7


class Datatype:
    def __init__(self, buf, off=0):
        b0 = buf[off]
        self.cls, self.version = b0 & 0x0F, b0 >> 4
        self.bits = buf[off + 1:off + 4]
        self.size = struct.unpack_from("<I", buf, off + 4)[0]
        self.props = off + 8
        if self.cls == 0:       # fixed point
            if self.bits[0] & 1:
                raise H5Error("big-endian integers not supported")
            signed = bool(self.bits[0] & 8)
            self.dtype = np.dtype(("<i" if signed else "<u") + str(self.size))
            self.nbytes = 8 + 4
        elif self.cls == 1:     # floating point
            if self.bits[0] & 1:
                raise H5Error("big-endian floats not supported")
            self.dtype = np.dtype("<f" + str(self.size))
            self.nbytes = 8 + 12
        elif self.cls == 3:     # fixed-length string
            self.dtype = np.dtype("S" + str(self.size))
            self.padding = self.bits[0] & 0x0F
            self.nbytes = 8
        elif self.cls == 6:     # compound, version 1, simple members: a numpy structured type
            if self.version != 1:
                raise H5Error(f"compound datatype version {self.version}")
            nmembers = struct.unpack_from("<H", buf, off + 1)[0]
            q, names, formats, offsets = off + 8, [], [], []
            for _ in range(nmembers):
                end = buf.index(b"\0", q)
                names.append(bytes(buf[q:end]).decode())
                q += _pad8(end - q + 1)
                offsets.append(struct.unpack_from("<I", buf, q)[0])
                if buf[q + 4] != 0:
                    raise H5Error("array members of compound datatypes not supported")
                q += 32
                m = Datatype(buf, q)
                formats.append(m.dtype)
                q += m.nbytes
            self.dtype = np.dtype({"names": names, "formats": formats, "offsets": offsets, "itemsize": self.size})
            self.nbytes = q - off
        else:
            raise H5Error(f"datatype class {self.cls} not supported")


class Dataspace:
    def __init__(self, buf, off=0):
        version, rank, flags = buf[off], buf[off + 1], buf[off + 2]
        if version == 1:
            p = off + 8
        elif version == 2:
            p = off + 4
        else:
            raise H5Error(f"dataspace version {version}")
        self.shape = tuple(struct.unpack_from("<Q", buf, p + 8 * k)[0] for k in range(rank))
        self.scalar = rank == 0
        self.nbytes = (p - off) + 8 * rank * (2 if flags & 1 else 1)


class Obj:
    """An object header: its messages, attributes and (for groups / datasets) the parsed essentials."""

    def __init__(self, f, addr):
        self.f, self.addr = f, addr
        self.attrs = {}
        self.attr_order = []
        self.messages = []
        self.symtab = None
        self.dtype = self.space = self.layout = None
        b = f.buf
        version, _, nmsg, _refcount, hsize = struct.unpack_from("<BBHII", b, addr)
        if version != 1:
            raise H5Error(f"object header version {version} at {addr}")
        blocks = [(addr + 16, hsize)]
        seen = 0
        while blocks and seen < nmsg:
            p, size = blocks.pop(0)
            end = p + size
            while p + 8 <= end and seen < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", b, p)
                body = p + 8
                self.messages.append((mtype, body, msize, mflags))
                seen += 1
                if mtype == 0x0010:
                    o, l = struct.unpack_from("<QQ", b, body)
                    blocks.append((o, l))
                p = body + msize
        if seen != nmsg:
            raise H5Error(f"object header at {addr}: found {seen} of {nmsg} messages")
        for mtype, body, msize, _ in self.messages:
            if mtype == 0x0011:
                self.symtab = struct.unpack_from("<QQ", b, body)
            elif mtype == 0x0001:
                self.space = Dataspace(b, body)
            elif mtype == 0x0003:
                self.dtype = Datatype(b, body)
            elif mtype == 0x0008:
                self.layout = self._layout(body)
            elif mtype == 0x000C:
                self._attribute(body)

    def _layout(self, p):
        b = self.f.buf
        version, cls = b[p], b[p + 1]
        if version != 3:
            raise H5Error(f"layout version {version}")
        if cls == 1:
            a, s = struct.unpack_from("<QQ", b, p + 2)
            return ("contiguous", a, s)
        if cls == 0:
            s = struct.unpack_from("<H", b, p + 2)[0]
            return ("compact", p + 4, s)
        if cls == 2:
            nd = b[p + 2]
            a = struct.unpack_from("<Q", b, p + 3)[0]
            dims = struct.unpack_from("<" + "I" * nd, b, p + 11)
            return ("chunked", a, dims)
        raise H5Error(f"layout class {cls}")

    def _attribute(self, p):
        b = self.f.buf
        version, _, nsize, tsize, ssize = struct.unpack_from("<BBHHH", b, p)
        if version != 1:
            raise H5Error(f"attribute version {version}")
        q = p + 8
        name = bytes(b[q:q + nsize]).split(b"\0")[0].decode()
        q += _pad8(nsize)
        dt = Datatype(b, q)
        q += _pad8(tsize)
        sp = Dataspace(b, q)
        q += _pad8(ssize)
        n = int(np.prod(sp.shape)) if not sp.scalar else 1
        a = np.frombuffer(b, dtype=dt.dtype, count=n, offset=q)
        if dt.cls == 3:
            v = [x.split(b"\0")[0].decode() for x in a.tolist()]
            v = v[0] if sp.scalar else v
        else:
            v = a[0].item() if sp.scalar else a.reshape(sp.shape).copy()
        self.attrs[name] = v
        self.attr_order.append(name)

    # ---- groups ----
    def links(self):
        if self.symtab is None:
            raise H5Error("not a group")
        btree, heap = self.symtab
        b = self.f.buf
        if bytes(b[heap:heap + 4]) != b"HEAP":
            raise H5Error("bad local heap signature")
        dseg_size, _free, dseg = struct.unpack_from("<QQQ", b, heap + 8)
        out = {}
        order = []
        node_keys = []

        def name_at(off):
            e = b.index(b"\0", dseg + off) if isinstance(b, (bytes, bytearray)) else None
            return bytes(b[dseg + off:e]).decode()

        def walk(node):
            if bytes(b[node:node + 4]) == b"SNOD":
                nsym = struct.unpack_from("<H", b, node + 6)[0]
                for k in range(nsym):
                    e = node + 8 + 40 * k
                    noff, oaddr = struct.unpack_from("<QQ", b, e)
                    nm = name_at(noff)
                    out[nm] = oaddr
                    order.append(nm)
                return
            if bytes(b[node:node + 4]) != b"TREE":
                raise H5Error(f"bad B-tree signature at {node}")
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            if ntype != 0:
                raise H5Error("not a group B-tree")
            p = node + 24
            keys = []
            for k in range(used):
                keys.append(struct.unpack_from("<Q", b, p)[0])
                child = struct.unpack_from("<Q", b, p + 8)[0]
                p += 16
                walk(child)
            keys.append(struct.unpack_from("<Q", b, p)[0])
            node_keys.append([name_at(k) for k in keys])

        if btree != UNDEF:
            walk(btree)
        self._link_order = order
        self.f.btree_keys[self.addr] = node_keys
        return out

    @property
    def link_order(self):
        self.links()
        return self._link_order

    def __getitem__(self, path):
        o = self
        for part in [p for p in path.split("/") if p]:
            l = o.links()
            if part not in l:
                raise KeyError(path)
            o = Obj(self.f, l[part])
        return o

    def __contains__(self, path):
        try:
            self[path]
            return True
        except KeyError:
            return False

    # ---- datasets ----
    def _filters(self):
        """ids of the filter pipeline (message 0x000B, version 1), in pipeline order"""
        b = self.f.buf
        for mtype, body, _, _ in self.messages:
            if mtype == 0x000B:
                n, q, ids = b[body + 1], body + 8, []
                for _k in range(n):
                    fid, nlen, _fl, ncd = struct.unpack_from("<HHHH", b, q)
                    ids.append(fid)
                    q += 8 + _pad8(nlen) + 4 * (ncd + (ncd & 1))
                return ids
        return []

    def read(self):
        if self.layout is None or self.dtype is None or self.space is None:
            raise H5Error("not a dataset")
        b = self.f.buf
        shape = self.space.shape
        n = int(np.prod(shape)) if shape else 1
        kind = self.layout[0]
        if kind in ("contiguous", "compact"):
            _, a, s = self.layout
            if kind == "contiguous" and a == UNDEF:   # never written: the library returns the fill value
                return np.zeros(shape, dtype=self.dtype.dtype)
            if s < n * self.dtype.size:
                raise H5Error("layout smaller than the dataspace")
            return np.frombuffer(b, dtype=self.dtype.dtype, count=n, offset=a).reshape(shape).copy()
        _, btree, cdims = self.layout
        nd = len(cdims) - 1
        out = np.zeros(shape, dtype=self.dtype.dtype)

        def walk(node):
            if bytes(b[node:node + 4]) != b"TREE":
                raise H5Error("bad chunk B-tree signature")
            ntype, level, used = struct.unpack_from("<BBH", b, node + 4)
            if ntype != 1:
                raise H5Error("not a chunk B-tree")
            ksize = 8 + 8 * (nd + 1)
            p = node + 24
            for k in range(used):
                csize, fmask = struct.unpack_from("<II", b, p)
                offs = struct.unpack_from("<" + "Q" * (nd + 1), b, p + 8)
                child = struct.unpack_from("<Q", b, p + ksize)[0]
                p += ksize + 8
                if level > 0:
                    walk(child)
                    continue
                raw = bytes(b[child:child + csize])
                for fid in reversed(self._filters()):
                    if fid == 1:
                        import zlib
                        raw = zlib.decompress(raw)
                    elif fid == 2:   # shuffle: byte b of element i sits at b * count + i
                        es = self.dtype.size
                        raw = np.frombuffer(raw, dtype=np.uint8).reshape(es, -1).T.tobytes()
                    else:
                        raise H5Error(f"filter {fid} not supported")
                if len(raw) != int(np.prod(cdims[:nd])) * self.dtype.size:
                    raise H5Error("unsupported chunk filter")
                c = np.frombuffer(raw, dtype=self.dtype.dtype).reshape(cdims[:nd])
                sl = tuple(slice(o, min(o + cd, sh)) for o, cd, sh in zip(offs[:nd], cdims[:nd], shape))
                out[sl] = c[tuple(slice(0, s.stop - s.start) for s in sl)]

        if btree != UNDEF:
            walk(btree)
        return out


class File(Obj):
    def __init__(self, path):
        self.buf = open(path, "rb").read()
        b = self.buf
        if b[:8] != b"\x89HDF\r\n\x1a\n":
            raise H5Error("not an HDF5 file")
        if b[8] != 0:
            raise H5Error(f"superblock version {b[8]} not supported")
        if b[13] != 8 or b[14] != 8:
            raise H5Error("only 8-byte offsets and lengths")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", b, 16)
        self.base, _fs, self.eof, _drv = struct.unpack_from("<QQQQ", b, 24)
        if self.eof != len(b):
            raise H5Error(f"end-of-file address {self.eof} != file size {len(b)}")
        self.root_entry = struct.unpack_from("<QQII", b, 56)
        self.btree_keys = {}
        super().__init__(self, self.root_entry[1])
