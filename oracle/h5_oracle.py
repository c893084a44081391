"""Independent pure-Python HDF5 reader -- TEST INFRASTRUCTURE ONLY (like everything under oracle/).

It follows the published HDF5 File Format Specification (version 1.1/2.0 structures that
libhdf5's default "earliest" format and h5py's default ``h5py.File(path, "w")`` produce, which is
what the reference writes: exp_mnist_resnet/save_kernel.py:26, cnn_gp/kernel_save_tools.py:21-23):
superblock version 0/1, symbol-table groups (version-1 B-tree of type 0, ``SNOD`` nodes, local
heap), version-1 object headers with continuation blocks, dataspace / datatype / fill-value /
layout messages, contiguous, compact and chunked (version-1 B-tree of type 1) storage.

Pinned against a file written by the real HDF5 library: tests/golden/libhdf5_matlab73.mat
(a copy of scipy's ``io/matlab/tests/data/testhdf5_7.4_GLNX86.mat``, written by MATLAB 7.4 through
libhdf5 behind a 512-byte user block); see tests/test_h5store.py.

The product writer/reader is cnn-gp_b200/csrc/h5store.cpp; this file shares no code with it and
is used to check the files it writes.
"""
import struct

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5FormatError(Exception):
    pass


class OracleDataset:
    def __init__(self, f, name, msgs):
        self.f, self.name = f, name
        self.shape = self.maxshape = None
        self.dtype = None
        self.fillvalue = None
        self.fill_defined = False
        self.layout = None      # ("contiguous", addr, size) | ("compact", bytes) | ("chunked", btree, chunk_dims)
        self.chunks = None
        self.filters = False
        self.attrs = {}
        for mtype, body in msgs:
            if mtype == 0x0001:
                self._dataspace(body)
            elif mtype == 0x0003:
                self.dtype = _datatype(body)
            elif mtype == 0x0005:
                self._fill(body)
            elif mtype == 0x0004 and not self.fill_defined:
                size, = struct.unpack_from("<I", body, 0)
                if size:
                    self._fill_raw = body[4:4 + size]
            elif mtype == 0x0008:
                self._layout(body)
            elif mtype == 0x000B:
                self.filters = True
        if self.dtype is not None and getattr(self, "_fill_raw", None) is not None:
            self.fillvalue = np.frombuffer(self._fill_raw, dtype=self.dtype, count=1)[0]
            self.fill_defined = True

    def _dataspace(self, b):
        ver, rank, flags = b[0], b[1], b[2]
        off = 8 if ver == 1 else 4
        dims = struct.unpack_from(f"<{rank}Q", b, off)
        off += 8 * rank
        mx = struct.unpack_from(f"<{rank}Q", b, off) if flags & 1 else dims
        self.shape = tuple(dims)
        self.maxshape = tuple(None if m == UNDEF else m for m in mx)

    def _fill(self, b):
        ver = b[0]
        if ver in (1, 2):
            defined = b[3]
            if ver == 1 or defined:
                size, = struct.unpack_from("<I", b, 4)
                self._fill_raw = b[8:8 + size] if size else None
        elif ver == 3:
            flags = b[1]
            if flags & 0x20:
                size, = struct.unpack_from("<I", b, 2)
                self._fill_raw = b[6:6 + size]
        else:
            raise H5FormatError(f"fill value message version {ver}")

    def _layout(self, b):
        ver, cls = b[0], b[1]
        if ver in (1, 2):  # libhdf5 <= 1.6
            nd, cls = b[1], b[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr, = struct.unpack_from("<Q", b, p)
                p += 8
            dims = struct.unpack_from(f"<{nd}I", b, p)
            p += 4 * nd
            if cls == 2:
                self.layout = ("chunked", addr, dims)
                self.chunks = tuple(dims[:-1])
            elif cls == 1:
                self.layout = ("contiguous", addr, None)
            else:
                size, = struct.unpack_from("<I", b, p)
                self.layout = ("compact", b[p + 4:p + 4 + size])
            return
        if ver != 3:
            raise H5FormatError(f"data layout message version {ver} (only 1-3 are read)")
        if cls == 0:
            size, = struct.unpack_from("<H", b, 2)
            self.layout = ("compact", b[4:4 + size])
        elif cls == 1:
            addr, size = struct.unpack_from("<QQ", b, 2)
            self.layout = ("contiguous", addr, size)
        elif cls == 2:
            nd = b[2]
            addr, = struct.unpack_from("<Q", b, 3)
            dims = struct.unpack_from(f"<{nd}I", b, 11)
            self.layout = ("chunked", addr, dims)
            self.chunks = tuple(dims[:-1])
        else:
            raise H5FormatError(f"layout class {cls}")

    # ------------------------------------------------------------------------------------
    def chunk_index(self):
        """{chunk offset tuple: (file address, stored bytes)} from the version-1 B-tree."""
        kind, addr, dims = self.layout
        assert kind == "chunked"
        out = {}
        if addr != UNDEF:
            self._walk(addr, len(dims), out, None)
        return out

    def _walk(self, addr, nd, out, expect_level):
        f = self.f
        hdr = f.at(addr, 24)
        if hdr[:4] != b"TREE":
            raise H5FormatError(f"no TREE signature at {addr}")
        ntype, level, used = hdr[4], hdr[5], struct.unpack_from("<H", hdr, 6)[0]
        if ntype != 1:
            raise H5FormatError("chunk index B-tree node is not of type 1")
        if expect_level is not None and level != expect_level:
            raise H5FormatError(f"B-tree level {level}, expected {expect_level}")
        ksz = 8 + 8 * nd
        body = f.at(addr + 24, used * (ksz + 8) + ksz)
        keys = []
        for e in range(used + 1):
            size, mask = struct.unpack_from("<II", body, e * (ksz + 8))
            offs = struct.unpack_from(f"<{nd}Q", body, e * (ksz + 8) + 8)
            keys.append((size, mask, offs))
        for e in range(used):
            child, = struct.unpack_from("<Q", body, e * (ksz + 8) + ksz)
            if not (keys[e][2] < keys[e + 1][2]):
                raise H5FormatError(f"B-tree keys not increasing at {addr}: {keys[e][2]} !< {keys[e+1][2]}")
            if level == 0:
                size, mask, offs = keys[e]
                if mask:
                    raise H5FormatError("filtered chunk")
                out[offs[:-1]] = (child, size)
            else:
                n0 = len(out)
                self._walk(child, nd, out, level - 1)
                lo = min(k for k in list(out)[n0:])
                if lo + (0,) < keys[e][2]:
                    raise H5FormatError("child chunk below its left key")
        return keys

    def read(self):
        """The whole dataset as a numpy array (fill value where nothing is stored)."""
        if self.filters:
            raise H5FormatError("filtered datasets are not read")
        kind = self.layout[0]
        n = int(np.prod(self.shape)) if self.shape else 1
        if kind == "compact":
            return np.frombuffer(self.layout[1], dtype=self.dtype, count=n).reshape(self.shape).copy()
        if kind == "contiguous":
            _, addr, size = self.layout
            if addr == UNDEF:
                return np.full(self.shape, self.fillvalue if self.fill_defined else 0, dtype=self.dtype)
            return np.frombuffer(self.f.at(addr, n * self.dtype.itemsize), dtype=self.dtype).reshape(self.shape).copy()
        out = np.full(self.shape, self.fillvalue if self.fill_defined else 0, dtype=self.dtype)
        cd = self.chunks
        nbytes = int(np.prod(cd)) * self.dtype.itemsize
        for offs, (addr, size) in self.chunk_index().items():
            if size != nbytes:
                raise H5FormatError(f"chunk of {size} bytes, expected {nbytes}")
            if any(o % c for o, c in zip(offs, cd)):
                raise H5FormatError(f"chunk offset {offs} not a multiple of the chunk shape")
            block = np.frombuffer(self.f.at(addr, nbytes), dtype=self.dtype).reshape(cd)
            sel = tuple(slice(o, min(o + c, s)) for o, c, s in zip(offs, cd, self.shape))
            if any(s.start >= s.stop for s in sel):
                continue  # chunk beyond the current extent
            out[sel] = block[tuple(slice(0, s.stop - s.start) for s in sel)]
        return out


def _datatype(b):
    cls, ver = b[0] & 0x0F, b[0] >> 4
    bits0, bits1 = b[1], b[2]
    size, = struct.unpack_from("<I", b, 4)
    order = ">" if bits0 & 1 else "<"
    if cls == 0:
        return np.dtype(f"{order}{'i' if bits0 & 8 else 'u'}{size}")
    if cls == 1:
        boff, prec, eloc, esize, mloc, msize, bias = struct.unpack_from("<HHBBBBI", b, 8)
        ieee = {4: (32, 23, 8, 0, 23, 127, 31), 8: (64, 52, 11, 0, 52, 1023, 63)}.get(size)
        if ieee is None or (prec, eloc, esize, mloc, msize, bias, bits1) != ieee or boff != 0:
            raise H5FormatError("floating-point type is not IEEE binary32/binary64")
        if (bits0 >> 4) & 3 != 2:
            raise H5FormatError("mantissa normalisation is not 'implied msb'")
        return np.dtype(f"{order}f{size}")
    if cls == 3:
        return np.dtype(f"S{size}")
    return np.dtype(f"V{size}")  # compound / reference / ...: opaque bytes of the right size


class OracleFile:
    def __init__(self, path):
        with open(path, "rb") as fh:
            self.data = fh.read()
        base = 0
        while self.data[base:base + 8] != SIGNATURE:  # user block: 0, 512, 1024, ...
            base = 512 if base == 0 else base * 2
            if base >= len(self.data):
                raise H5FormatError("no HDF5 signature")
        self.sb_offset = base
        d = self.data
        self.sb_version = d[base + 8]
        if self.sb_version > 1:
            raise H5FormatError(f"superblock version {self.sb_version} (only 0 and 1 are read)")
        so, sl = d[base + 13], d[base + 14]
        if (so, sl) != (8, 8):
            raise H5FormatError("offsets/lengths are not 8 bytes")
        self.leaf_k, self.internal_k = struct.unpack_from("<HH", d, base + 16)
        p = base + 24
        self.chunk_k = 32
        if self.sb_version == 1:
            self.chunk_k, = struct.unpack_from("<H", d, p)
            p += 4
        self.base, self.freespace, self.eof, self.driver = struct.unpack_from("<QQQQ", d, p)
        p += 32
        self.root_entry = self._entry(d[p:p + 40])
        if self.eof > len(self.data):
            raise H5FormatError("end-of-file address beyond the file (truncated)")
        self.eof_matches_size = self.eof == len(self.data)  # true for every file libhdf5 closed cleanly
        self.datasets = {}
        self.groups = []
        self._group(self.root_entry, "")

    def at(self, addr, n):
        a = self.base + addr
        if addr == UNDEF or a + n > len(self.data):
            raise H5FormatError(f"address {addr} (+{n}) outside the file")
        return self.data[a:a + n]

    @staticmethod
    def _entry(b):
        name_off, ohdr, ctype = struct.unpack_from("<QQI", b, 0)
        return dict(name_off=name_off, ohdr=ohdr, ctype=ctype, scratch=b[24:40])

    def _messages(self, addr):
        hdr = self.at(addr, 16)
        ver, nmsg = hdr[0], struct.unpack_from("<H", hdr, 2)[0]
        if ver != 1:
            raise H5FormatError(f"object header version {ver} at {addr}")
        refcount, hsize = struct.unpack_from("<II", hdr, 4)
        # what libhdf5 insists on when it loads a version-1 header (H5Ocache.c): chunk and message
        # sizes are multiples of 8, messages start 8-byte aligned, and the message count in the
        # prefix is exactly the number of messages found (NIL and continuation messages included)
        if hsize % 8 or addr % 8:
            raise H5FormatError(f"object header at {addr}: size {hsize} / address not 8-byte aligned")
        blocks = [(addr + 16, hsize)]
        msgs = []
        while blocks:
            a, n = blocks.pop(0)
            blk = self.at(a, n)
            p = 0
            while p + 8 <= n and len(msgs) < nmsg:
                mtype, msize, mflags = struct.unpack_from("<HHB", blk, p)
                if msize % 8 or p + 8 + msize > n:
                    raise H5FormatError(f"object header at {addr}: message 0x{mtype:04x} of {msize} bytes is "
                                        "misaligned or overruns its block")
                body = blk[p + 8:p + 8 + msize]
                p += 8 + msize
                msgs.append((mtype, body))
                if mtype == 0x0010:
                    ca, cn = struct.unpack_from("<QQ", body, 0)
                    blocks.append((ca, cn))
            if len(msgs) < nmsg and not blocks and p + 8 <= n:
                raise H5FormatError(f"object header at {addr}: trailing bytes without a message header")
        if len(msgs) != nmsg:
            raise H5FormatError(f"object header at {addr}: prefix says {nmsg} messages, found {len(msgs)}")
        if refcount < 1:
            raise H5FormatError(f"object header at {addr}: reference count {refcount}")
        return [(t, b) for t, b in msgs if t not in (0x0000, 0x0010)]

    def _group(self, entry, prefix):
        msgs = self._messages(entry["ohdr"])
        stab = [b for t, b in msgs if t == 0x0011]
        if not stab:
            raise H5FormatError("group without a symbol-table message (new-style groups are not read)")
        btree, heap = struct.unpack_from("<QQ", stab[0], 0)
        if entry["ctype"] == 1:
            cb, ch = struct.unpack_from("<QQ", entry["scratch"], 0)
            if (cb, ch) != (btree, heap):
                raise H5FormatError("cached B-tree/heap addresses differ from the symbol-table message")
        h = self.at(heap, 32)
        if h[:4] != b"HEAP":
            raise H5FormatError("no HEAP signature")
        hsize, hfree, hdata = struct.unpack_from("<QQQ", h, 8)
        if hfree != 1 and hfree >= hsize:
            raise H5FormatError("local heap free-list head outside the data segment")
        seg = self.at(hdata, hsize)
        # walk the free list like libhdf5 does when it loads a heap
        fl = hfree
        while fl != 1:
            if fl % 8 or fl + 16 > hsize:
                raise H5FormatError("bad local heap free block")
            nxt, sz = struct.unpack_from("<QQ", seg, fl)
            if sz < 16 or fl + sz > hsize:
                raise H5FormatError("bad local heap free block size")
            fl = nxt
        self.groups.append(dict(path=prefix or "/", btree=btree, heap=heap, heap_size=hsize))
        names = []
        self._gnode(btree, seg, prefix, names)
        if names != sorted(names):
            raise H5FormatError(f"link names are not in B-tree order: {names}")

    def _name(self, seg, off):
        end = seg.index(b"\0", off)
        return seg[off:end].decode()

    def _gnode(self, addr, seg, prefix, names):
        hdr = self.at(addr, 24)
        if hdr[:4] != b"TREE" or hdr[4] != 0:
            raise H5FormatError("group B-tree node expected")
        level, used = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
        if used > 2 * self.internal_k:
            raise H5FormatError("group B-tree node over capacity")
        body = self.at(addr + 24, used * 16 + 8)
        for e in range(used):
            k0, child, k1 = struct.unpack_from("<QQQ", body, e * 16)
            if level > 0:
                self._gnode(child, seg, prefix, names)
                continue
            sn = self.at(child, 8)
            if sn[:4] != b"SNOD" or sn[4] != 1:
                raise H5FormatError("no SNOD signature")
            n, = struct.unpack_from("<H", sn, 6)
            if n > 2 * self.leaf_k:
                raise H5FormatError("symbol-table node over capacity")
            ents = self.at(child + 8, 40 * n)
            lo, hi = self._name(seg, k0), self._name(seg, k1)
            for i in range(n):
                ent = self._entry(ents[40 * i:40 * i + 40])
                name = self._name(seg, ent["name_off"])
                if not (lo < name <= hi) and not (lo == "" and name <= hi):
                    raise H5FormatError(f"link {name!r} outside its B-tree key range ({lo!r}, {hi!r}]")
                names.append(name)
                path = prefix + "/" + name
                if ent["ctype"] == 1:
                    self._group(ent, path)
                    continue
                msgs = self._messages(ent["ohdr"])
                if any(t == 0x0011 for t, _ in msgs):
                    self._group(ent, path)
                else:
                    self.datasets[path.lstrip("/")] = OracleDataset(self, path, msgs)
