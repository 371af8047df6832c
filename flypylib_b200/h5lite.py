"""Minimal HDF5 reader / writer (h5py is not installed on the target image).

What the hot path needs from HDF5, and nothing else:
  * Keras ``.keras.h5`` / ``.h5`` weight files: ``load_network`` / ``load_model`` (flypylib/fplnetwork.py:32-44,
    81-97): group ``model_weights`` (or the file root for ``save_weights`` files), attribute ``layer_names``, per
    layer attribute ``weight_names`` and one float32 dataset per weight;
  * ``/main`` volumes given as a path to ``FplNetwork.infer`` (fplnetwork.py:137-139) and ``voxel2obj``
    (fplobjdetect.py:154-156, 161-163): contiguous or chunked (gzip / shuffle) datasets.

Reader: superblock versions 0-3, object headers version 1 and 2, old-style groups (symbol table: B-tree v1 + local
heap) and new-style groups with compact link messages, dataspace v1/v2, datatypes fixed-point / floating-point /
string (fixed and variable length through the global heap), data layout v1-v3 (compact, contiguous, chunked with the
deflate and shuffle filters), attributes v1-v3, header continuation blocks.  Dense link / attribute storage (fractal
heaps), layout v4, compound and reference types raise ``H5Error``.

Writer (``write_h5``): what h5py's default (``libver='earliest'``) produces for plain trees of groups, contiguous
little-endian datasets and attributes: superblock 0, version-1 object headers, symbol-table groups.

Format: "HDF5 File Format Specification Version 3.0" (The HDF Group).
"""
import struct
import zlib

import numpy as np

SIGNATURE = b"\x89HDF\r\n\x1a\n"
UNDEF = 0xFFFFFFFFFFFFFFFF


class H5Error(Exception):
    pass


# ------------------------------------------------------------------------------------------------
# reader
# ------------------------------------------------------------------------------------------------
class _Datatype(object):
    def __init__(self, cls, size, dtype=None, vlen_string=False, base=None, pad=0):
        self.cls, self.size, self.dtype, self.vlen_string, self.base, self.pad = cls, size, dtype, vlen_string, base, pad


class Dataset(object):
    def __init__(self, f, name, shape, dt, layout, filters, attrs):
        self._f, self.name, self.shape, self._dt, self._layout, self._filters, self.attrs = f, name, shape, dt, layout, filters, attrs

    @property
    def dtype(self):
        return self._dt.dtype if self._dt.dtype is not None else np.dtype(object)

    def __getitem__(self, key):
        return self.read()[key]

    def read(self):
        f, dt = self._f, self._dt
        n = int(np.prod(self.shape)) if self.shape else 1
        kind = self._layout[0]
        if kind == "compact":
            raw = self._layout[1]
        elif kind == "contiguous":
            addr, size = self._layout[1], self._layout[2]
            raw = b"\0" * (n * dt.size) if addr == UNDEF else f._read(addr, n * dt.size)
        else:
            raw = self._read_chunked(n)
        return f._decode(raw, dt, self.shape)

    def _read_chunked(self, n):
        f, dt = self._f, self._dt
        _, btree, cdims = self._layout
        rank = len(self.shape)
        chunk = tuple(cdims[:rank])
        out = np.zeros(self.shape, dtype=np.uint8 if dt.dtype is None else dt.dtype)
        if dt.dtype is None:
            raise H5Error("chunked variable-length data is not supported")
        if btree == UNDEF:
            return out.tobytes()
        for offs, size, mask, addr in f._chunk_index(btree, rank):
            raw = f._read(addr, size)
            for i, (fid, cd) in reversed(list(enumerate(self._filters))):
                if mask & (1 << i):
                    continue
                if fid == 1:
                    raw = zlib.decompress(raw)
                elif fid == 2:
                    es = cd[0] if cd else dt.size
                    a = np.frombuffer(raw, dtype=np.uint8)
                    m = a.size // es
                    raw = a[:m * es].reshape(es, m).T.tobytes() + a[m * es:].tobytes()
                elif fid == 3:
                    raw = raw[:-4]                   # fletcher32 checksum: not verified
                else:
                    raise H5Error("unsupported filter id %d" % fid)
            block = np.frombuffer(raw, dtype=dt.dtype, count=int(np.prod(chunk))).reshape(chunk)
            sl_out, sl_in = [], []
            for o, c, s in zip(offs, chunk, self.shape):
                e = min(o + c, s)
                sl_out.append(slice(o, e)); sl_in.append(slice(0, e - o))
            out[tuple(sl_out)] = block[tuple(sl_in)]
        return out.tobytes()


class Group(object):
    def __init__(self, f, name, links, attrs):
        self._f, self.name, self._links, self.attrs = f, name, links, attrs

    def keys(self):
        return list(self._links)

    def __contains__(self, name):
        try:
            self[name]
            return True
        except KeyError:
            return False

    def __iter__(self):
        return iter(self._links)

    def __getitem__(self, path):
        node = self
        parts = [p for p in path.split("/") if p]
        if path.startswith("/"):
            node = self._f.root
        for p in parts:
            if not isinstance(node, Group) or p not in node._links:
                raise KeyError(path)
            base = node.name.rstrip("/")
            node = node._f._object(node._links[p], base + "/" + p)
        return node


class File(object):
    """Read-only HDF5 file.  ``File(path)['/main'][:]``, ``.attrs``, ``.keys()`` as with h5py."""

    def __init__(self, path):
        with open(path, "rb") as fh:
            self._buf = fh.read()
        self._cache = {}
        base = 0
        while True:
            if self._buf[base:base + 8] == SIGNATURE:
                break
            base = 512 if base == 0 else base * 2
            if base + 8 > len(self._buf):
                raise H5Error("%s: not an HDF5 file (no superblock signature)" % path)
        b = self._buf
        ver = b[base + 8]
        if ver in (0, 1):
            self.O, self.L = b[base + 13], b[base + 14]
            p = base + 24 + (4 if ver == 1 else 0)
            self.base = self._uint(p, self.O, raw=True)
            p += 4 * self.O                                   # base, free-space, end-of-file, driver-info addresses
            root_addr = self._uint(p + self.O, self.O, raw=True)       # root symbol-table entry: name offset, header address
        elif ver in (2, 3):
            self.O, self.L = b[base + 9], b[base + 10]
            p = base + 12
            self.base = self._uint(p, self.O, raw=True)
            root_addr = self._uint(p + 3 * self.O, self.O, raw=True)
        else:
            raise H5Error("unsupported superblock version %d" % ver)
        if self.base == UNDEF or (self.base == 0 and base):
            self.base = base              # user block in front of the superblock: addresses are relative to it
        self.root = self._object(root_addr, "/")

    # ---- h5py-like surface
    def __getitem__(self, path):
        return self.root[path]

    def __contains__(self, path):
        return path in self.root

    def keys(self):
        return self.root.keys()

    @property
    def attrs(self):
        return self.root.attrs

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    # ---- low level
    def _uint(self, pos, size, raw=False):
        if not raw:
            pos += self.base
        v = int.from_bytes(self._buf[pos:pos + size], "little")
        if size == 8 and v == UNDEF or size == 4 and v == 0xFFFFFFFF:
            return UNDEF
        return v

    def _read(self, addr, size):
        a = addr + self.base
        if a + size > len(self._buf):
            raise H5Error("read beyond the end of the file (truncated file?)")
        return self._buf[a:a + size]

    def _messages(self, addr):
        """-> list of (type, flags, bytes) of the object header at addr (versions 1 and 2, continuations followed)."""
        b, out = self._buf, []
        a = addr + self.base
        if b[a:a + 4] == b"OHDR":
            flags = b[a + 5]
            p = a + 6
            if flags & 0x20:
                p += 16
            if flags & 0x10:
                p += 4
            szlen = 1 << (flags & 3)
            size0 = int.from_bytes(b[p:p + szlen], "little")
            p += szlen
            blocks = [(p, p + size0)]
            track = bool(flags & 0x04)
            while blocks:
                p, end = blocks.pop(0)
                while p + 4 <= end:
                    mtype, msize, mflags = b[p], int.from_bytes(b[p + 1:p + 3], "little"), b[p + 3]
                    p += 4 + (2 if track else 0)
                    data = b[p:p + msize]
                    p += msize
                    if mtype == 0x10:
                        off = int.from_bytes(data[:self.O], "little") + self.base
                        ln = int.from_bytes(data[self.O:self.O + self.L], "little")
                        blocks.append((off + 4, off + ln - 4))      # "OCHK" ... checksum
                    elif mtype != 0:
                        out.append((mtype, mflags, data))
            return out
        if b[a] != 1:
            raise H5Error("unsupported object header version %d at %d" % (b[a], addr))
        nmsg = int.from_bytes(b[a + 2:a + 4], "little")
        size = int.from_bytes(b[a + 8:a + 12], "little")
        blocks = [(a + 16, a + 16 + size)]
        while blocks and len(out) < nmsg + 64:
            p, end = blocks.pop(0)
            while p + 8 <= end:
                mtype = int.from_bytes(b[p:p + 2], "little")
                msize = int.from_bytes(b[p + 2:p + 4], "little")
                mflags = b[p + 4]
                data = b[p + 8:p + 8 + msize]
                p += 8 + msize
                if mtype == 0x10:
                    off = int.from_bytes(data[:self.O], "little") + self.base
                    ln = int.from_bytes(data[self.O:self.O + self.L], "little")
                    blocks.append((off, off + ln))
                elif mtype != 0:
                    out.append((mtype, mflags, data))
        return out

    def _datatype(self, d, pos=0):
        """-> (_Datatype, bytes consumed)"""
        cls, ver = d[pos] & 0x0F, d[pos] >> 4
        bits = d[pos + 1] | (d[pos + 2] << 8) | (d[pos + 3] << 16)
        size = int.from_bytes(d[pos + 4:pos + 8], "little")
        order = ">" if bits & 1 else "<"
        if cls == 0:
            return _Datatype(0, size, np.dtype("%s%s%d" % (order, "i" if bits & 8 else "u", size))), 12
        if cls == 1:
            return _Datatype(1, size, np.dtype("%sf%d" % (order, size))), 20
        if cls == 3:
            return _Datatype(3, size, np.dtype("S%d" % size), pad=bits & 0x0F), 8
        if cls == 9:
            base, used = self._datatype(d, pos + 8)
            return _Datatype(9, size, None, vlen_string=(bits & 0x0F) == 1, base=base), 8 + used
        if cls == 8:        # enumeration (h5py stores bool as an int8 enum): read as the base integer
            base, _ = self._datatype(d, pos + 8)
            return _Datatype(0, size, base.dtype), 0
        raise H5Error("unsupported datatype class %d (version %d)" % (cls, ver))

    def _dataspace(self, d):
        ver, rank, flags = d[0], d[1], d[2]
        if ver == 1:
            p = 8
        elif ver == 2:
            if d[3] == 2:
                return None                   # null dataspace
            p = 4
        else:
            raise H5Error("unsupported dataspace version %d" % ver)
        return tuple(int.from_bytes(d[p + i * self.L:p + (i + 1) * self.L], "little") for i in range(rank))

    def _decode(self, raw, dt, shape):
        n = int(np.prod(shape)) if shape else 1
        if dt.cls == 9:
            if not dt.vlen_string and dt.base.dtype is None:
                raise H5Error("nested variable-length data is not supported")
            out = []
            step = 4 + self.O + 4
            for i in range(n):
                ln = int.from_bytes(raw[i * step:i * step + 4], "little")
                coll = int.from_bytes(raw[i * step + 4:i * step + 4 + self.O], "little")
                idx = int.from_bytes(raw[i * step + 4 + self.O:i * step + step], "little")
                data = self._global_heap(coll, idx) if ln or coll not in (0, UNDEF) else b""
                if dt.vlen_string:
                    out.append(data[:ln].decode("utf-8", "replace"))
                else:
                    out.append(np.frombuffer(data, dtype=dt.base.dtype, count=ln).copy())
            if not shape:
                return out[0]
            arr = np.empty(n, dtype=object)
            arr[:] = out
            return arr.reshape(shape)
        arr = np.frombuffer(raw, dtype=dt.dtype, count=n)
        if dt.cls == 3:
            if not shape:
                return bytes(arr[0]).rstrip(b"\0 ") if dt.pad != 2 else bytes(arr[0]).rstrip(b" ")
            return arr.reshape(shape).copy()
        arr = arr.reshape(shape).copy() if shape else arr[0]
        if isinstance(arr, np.ndarray) and arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("="))
        return arr

    def _global_heap(self, coll, idx):
        a = coll + self.base
        b = self._buf
        if b[a:a + 4] != b"GCOL":
            raise H5Error("bad global heap collection at %d" % coll)
        size = int.from_bytes(b[a + 8:a + 8 + self.L], "little")
        p, end = a + 8 + self.L, a + size
        while p + 8 + self.L <= end:
            oi = int.from_bytes(b[p:p + 2], "little")
            osz = int.from_bytes(b[p + 8:p + 8 + self.L], "little")
            if oi == idx:
                return b[p + 8 + self.L:p + 8 + self.L + osz]
            if oi == 0:
                break
            p += 8 + self.L + ((osz + 7) // 8) * 8
        raise H5Error("global heap object %d not found" % idx)

    def _attribute(self, d):
        ver = d[0]
        nsz, tsz, ssz = (int.from_bytes(d[2 + 2 * i:4 + 2 * i], "little") for i in range(3))
        if ver == 1:
            pad = lambda v: (v + 7) // 8 * 8        # noqa: E731
            p = 8
        elif ver in (2, 3):
            if d[1] & 3:
                raise H5Error("shared attribute datatypes / dataspaces are not supported")
            pad = lambda v: v                        # noqa: E731
            p = 8 if ver == 2 else 9
        else:
            raise H5Error("unsupported attribute version %d" % ver)
        name = d[p:p + nsz].split(b"\0")[0].decode("utf-8")
        p += pad(nsz)
        dt, _ = self._datatype(d, p)
        p += pad(tsz)
        shape = self._dataspace(d[p:p + ssz])
        p += pad(ssz)
        if shape is None:
            return name, None
        return name, self._decode(d[p:], dt, shape)

    def _symbol_table(self, btree, heap):
        b = self._buf
        h = heap + self.base
        if b[h:h + 4] != b"HEAP":
            raise H5Error("bad local heap at %d" % heap)
        data_addr = int.from_bytes(b[h + 8 + 2 * self.L:h + 8 + 2 * self.L + self.O], "little") + self.base
        links = {}

        def walk(addr):
            a = addr + self.base
            if b[a:a + 4] == b"SNOD":
                n = int.from_bytes(b[a + 6:a + 8], "little")
                p = a + 8
                for _ in range(n):
                    noff = int.from_bytes(b[p:p + self.O], "little")
                    oaddr = int.from_bytes(b[p + self.O:p + 2 * self.O], "little")
                    s = data_addr + noff
                    name = b[s:b.index(b"\0", s)].decode("utf-8")
                    links[name] = oaddr
                    p += 2 * self.O + 24
                return
            if b[a:a + 4] != b"TREE":
                raise H5Error("bad B-tree node at %d" % addr)
            n = int.from_bytes(b[a + 6:a + 8], "little")
            p = a + 8 + 2 * self.O + self.L               # first child (after key 0)
            for _ in range(n):
                walk(int.from_bytes(b[p:p + self.O], "little"))
                p += self.O + self.L
        if btree != UNDEF:
            walk(btree)
        return links

    def _chunk_index(self, addr, rank):
        b = self._buf
        a = addr + self.base
        if b[a:a + 4] != b"TREE":
            raise H5Error("bad chunk B-tree node at %d" % addr)
        level = b[a + 5]
        n = int.from_bytes(b[a + 6:a + 8], "little")
        p = a + 8 + 2 * self.O
        ksz = 8 + 8 * (rank + 1)
        for _ in range(n):
            size = int.from_bytes(b[p:p + 4], "little")
            mask = int.from_bytes(b[p + 4:p + 8], "little")
            offs = tuple(int.from_bytes(b[p + 8 + 8 * i:p + 16 + 8 * i], "little") for i in range(rank))
            child = int.from_bytes(b[p + ksz:p + ksz + self.O], "little")
            if level == 0:
                yield offs, size, mask, child
            else:
                for item in self._chunk_index(child, rank):
                    yield item
            p += ksz + self.O

    def _object(self, addr, name):
        if addr in self._cache:
            return self._cache[addr]
        attrs, links = {}, {}
        shape = dt = layout = None
        filters = []
        is_group = False
        for mtype, mflags, d in self._messages(addr):
            if mflags & 2 and mtype in (1, 3, 0x0B):
                raise H5Error("shared header messages are not supported")
            if mtype == 0x11:
                is_group = True
                links.update(self._symbol_table(int.from_bytes(d[:self.O], "little"),
                                                int.from_bytes(d[self.O:2 * self.O], "little")))
            elif mtype == 0x02:
                is_group = True
                p = 2 + (8 if d[1] & 1 else 0)
                fheap = int.from_bytes(d[p:p + self.O], "little")
                if fheap != UNDEF and (self.O != 8 or fheap != 0xFFFFFFFFFFFFFFFF):
                    raise H5Error("%s: dense link storage (fractal heap) is not supported" % name)
            elif mtype == 0x06:
                is_group = True
                fl = d[1]
                p = 2
                ltype = 0
                if fl & 8:
                    ltype = d[p]; p += 1
                if fl & 4:
                    p += 8
                if fl & 16:
                    p += 1
                ll = 1 << (fl & 3)
                nlen = int.from_bytes(d[p:p + ll], "little"); p += ll
                lname = d[p:p + nlen].decode("utf-8"); p += nlen
                if ltype == 0:
                    links[lname] = int.from_bytes(d[p:p + self.O], "little")
            elif mtype == 0x01:
                shape = self._dataspace(d)
            elif mtype == 0x03:
                dt, _ = self._datatype(d)
            elif mtype == 0x08:
                layout = self._layout(d)
            elif mtype == 0x0B:
                filters = self._filter_pipeline(d)
            elif mtype == 0x0C:
                k, v = self._attribute(d)
                attrs[k] = v
            elif mtype == 0x15 and len(d) > 2:
                p = 2 + (2 if d[1] & 1 else 0)
                if int.from_bytes(d[p:p + self.O], "little") not in (UNDEF,):
                    raise H5Error("%s: dense attribute storage (fractal heap) is not supported" % name)
        if dt is not None and layout is not None:
            obj = Dataset(self, name, shape if shape is not None else (), dt, layout, filters, attrs)
        elif is_group or name == "/":
            obj = Group(self, name, links, attrs)
        else:
            raise H5Error("%s: object is neither a group nor a dataset this reader understands" % name)
        self._cache[addr] = obj
        return obj

    def _layout(self, d):
        ver = d[0]
        if ver == 3:
            cls = d[1]
            if cls == 0:
                size = int.from_bytes(d[2:4], "little")
                return ("compact", bytes(d[4:4 + size]))
            if cls == 1:
                return ("contiguous", self._u(d, 2, self.O), self._u(d, 2 + self.O, self.L))
            if cls == 2:
                rank = d[2]
                bt = self._u(d, 3, self.O)
                dims = [int.from_bytes(d[3 + self.O + 4 * i:7 + self.O + 4 * i], "little") for i in range(rank)]
                return ("chunked", bt, dims)
        elif ver in (1, 2):
            rank, cls = d[1], d[2]
            p = 8
            addr = UNDEF
            if cls != 0:
                addr = self._u(d, p, self.O); p += self.O
            dims = [int.from_bytes(d[p + 4 * i:p + 4 * i + 4], "little") for i in range(rank)]
            p += 4 * rank
            if cls == 1:
                return ("contiguous", addr, 0)
            if cls == 2:
                return ("chunked", addr, dims)
            size = int.from_bytes(d[p:p + 4], "little")
            return ("compact", bytes(d[p + 4:p + 4 + size]))
        raise H5Error("unsupported data layout message (version %d)" % ver)

    def _u(self, d, p, size):
        v = int.from_bytes(d[p:p + size], "little")
        return UNDEF if v == (1 << (8 * size)) - 1 else v

    def _filter_pipeline(self, d):
        ver, n = d[0], d[1]
        p = 8 if ver == 1 else 2
        out = []
        for _ in range(n):
            fid = int.from_bytes(d[p:p + 2], "little"); p += 2
            nlen = 0
            if ver == 1 or fid >= 256:
                nlen = int.from_bytes(d[p:p + 2], "little"); p += 2
            p += 2                                              # flags
            ncd = int.from_bytes(d[p:p + 2], "little"); p += 2
            p += (nlen + 7) // 8 * 8 if ver == 1 else nlen
            cd = [int.from_bytes(d[p + 4 * i:p + 4 * i + 4], "little") for i in range(ncd)]
            p += 4 * ncd
            if ver == 1 and ncd % 2:
                p += 4
            out.append((fid, cd))
        return out


# ------------------------------------------------------------------------------------------------
# writer
# ------------------------------------------------------------------------------------------------
def _pad8(b):
    return b + b"\0" * (-len(b) % 8)


def _dtype_msg(dt):
    dt = np.dtype(dt)
    if dt.kind == "f":
        props = {4: (31, 23, 8, 0, 23, 127), 8: (63, 52, 11, 0, 52, 1023), 2: (15, 10, 5, 0, 10, 15)}[dt.itemsize]
        sign, eloc, esz, mloc, msz, bias = props
        bits = 0x20 | (sign << 8)                 # little-endian, mantissa normalisation 2 (implied msb), sign location
        return struct.pack("<BBBBI", 0x11, bits & 0xFF, (bits >> 8) & 0xFF, 0, dt.itemsize) + \
            struct.pack("<HHBBBBI", 0, dt.itemsize * 8, eloc, esz, mloc, msz, bias)
    if dt.kind in "iu":
        bits = 0x08 if dt.kind == "i" else 0
        return struct.pack("<BBBBI", 0x10, bits, 0, 0, dt.itemsize) + struct.pack("<HH", 0, dt.itemsize * 8)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x13, 0x01, 0, 0, dt.itemsize)       # null-padded ASCII
    raise H5Error("write_h5: unsupported dtype %s" % dt)


def _space_msg(shape):
    return struct.pack("<BBBB4x", 1, len(shape), 0, 0) + b"".join(struct.pack("<Q", s) for s in shape)


def _attr_msg(name, value):
    if isinstance(value, str):
        value = value.encode("utf-8")
    if isinstance(value, bytes):
        value = np.array(value if value else b"\0", dtype="S%d" % max(1, len(value)))
    value = np.asarray(value)
    if value.dtype.kind == "U":
        value = np.char.encode(value, "utf-8")
    if value.dtype.byteorder == ">":
        value = value.astype(value.dtype.newbyteorder("<"))
    nm = name.encode("utf-8") + b"\0"
    dtm, spm = _dtype_msg(value.dtype), _space_msg(value.shape)
    return struct.pack("<BBHHH", 1, 0, len(nm), len(dtm), len(spm)) + _pad8(nm) + _pad8(dtm) + _pad8(spm) + \
        np.ascontiguousarray(value).tobytes()


def _header(msgs):
    body = b""
    for mtype, data in msgs:
        data = _pad8(data)
        body += struct.pack("<HHB3x", mtype, len(data), 0) + data
    return struct.pack("<BBHII4x", 1, 0, len(msgs), 1, len(body)) + body


class _Writer(object):
    LEAF_K, NODE_K = 64, 16

    def __init__(self):
        self.buf = bytearray()

    def alloc(self, data):
        self.buf += b"\0" * (-len(self.buf) % 8)
        addr = len(self.buf)
        self.buf += data
        return addr

    def dataset(self, arr, attrs, chunks=None, gzip=None, shuffle=False):
        arr = np.ascontiguousarray(arr)
        if arr.dtype.byteorder == ">":
            arr = arr.astype(arr.dtype.newbyteorder("<"))
        msgs = [(0x01, _space_msg(arr.shape)), (0x03, _dtype_msg(arr.dtype))]
        if chunks is None:
            raw = arr.tobytes()
            daddr = self.alloc(raw) if raw else UNDEF
            msgs.append((0x08, struct.pack("<BBQQ", 3, 1, daddr, len(raw))))
        else:
            # chunked layout (B-tree v1 with one leaf node), optional shuffle + deflate filters
            rank, es = arr.ndim, arr.dtype.itemsize
            grid = [range(0, s, c) for s, c in zip(arr.shape, chunks)]
            entries = []
            for offs in np.ndindex(*[len(g) for g in grid]):
                o = [grid[i][j] for i, j in enumerate(offs)]
                block = np.zeros(chunks, dtype=arr.dtype)
                sl = tuple(slice(a, min(a + c, s)) for a, c, s in zip(o, chunks, arr.shape))
                block[tuple(slice(0, x.stop - x.start) for x in sl)] = arr[sl]
                raw = block.tobytes()
                if shuffle:
                    raw = np.frombuffer(raw, dtype=np.uint8).reshape(-1, es).T.tobytes()
                if gzip is not None:
                    raw = zlib.compress(raw, gzip)
                entries.append((o, len(raw), self.alloc(raw)))
            if len(entries) > 64:
                raise H5Error("write_h5: more than 64 chunks")
            tree = b"TREE" + struct.pack("<BBHQQ", 1, 0, len(entries), UNDEF, UNDEF)
            for o, size, addr in entries:
                tree += struct.pack("<II", size, 0) + b"".join(struct.pack("<Q", v) for v in o + [0]) + struct.pack("<Q", addr)
            tree += struct.pack("<II", 0, 0) + b"".join(struct.pack("<Q", v) for v in list(arr.shape) + [0])
            tree += b"\0" * ((64 - len(entries)) * (8 + 8 * (rank + 1) + 8))
            taddr = self.alloc(tree)
            msgs.append((0x08, struct.pack("<BBB", 3, 2, rank + 1) + struct.pack("<Q", taddr) +
                         b"".join(struct.pack("<I", c) for c in list(chunks) + [es])))
            filt = []
            if shuffle:
                filt.append(struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I", es) + b"\0" * 4)
            if gzip is not None:
                filt.append(struct.pack("<HHHH", 1, 0, 0, 1) + struct.pack("<I", gzip) + b"\0" * 4)
            if filt:
                msgs.append((0x0B, struct.pack("<BB6x", 1, len(filt)) + b"".join(filt)))
        msgs += [(0x0C, _attr_msg(k, v)) for k, v in attrs.items()]
        return self.alloc(_header(msgs))

    def group(self, children, attrs):
        """children: {name: object header address} -> object header address"""
        names = sorted(children, key=lambda s: s.encode("utf-8"))
        if len(names) > 2 * self.LEAF_K:
            raise H5Error("write_h5: more than %d links in one group" % (2 * self.LEAF_K))
        heap = bytearray(b"\0" * 8)                       # offset 0: the empty name
        offs = []
        for nm in names:
            offs.append(len(heap))
            heap += _pad8(nm.encode("utf-8") + b"\0")
        heap_data = self.alloc(bytes(heap))
        heap_addr = self.alloc(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), UNDEF, heap_data))
        snod = b"SNOD" + struct.pack("<BBH", 1, 0, len(names))
        for nm, off in zip(names, offs):
            snod += struct.pack("<QQII16x", off, children[nm], 0, 0)
        snod += b"\0" * ((2 * self.LEAF_K - len(names)) * 40)
        snod_addr = self.alloc(snod)
        tree = b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, UNDEF, UNDEF)
        tree += struct.pack("<QQQ", 0, snod_addr, offs[-1] if offs else 0)
        tree += b"\0" * ((2 * self.NODE_K - 1) * 16)
        tree_addr = self.alloc(tree)
        msgs = [(0x11, struct.pack("<QQ", tree_addr, heap_addr))]
        msgs += [(0x0C, _attr_msg(k, v)) for k, v in attrs.items()]
        return self.alloc(_header(msgs)), tree_addr, heap_addr

    def node(self, tree):
        """tree: {'attrs': {...}, 'items': {name: ndarray | subtree}} -> (header address, btree, heap)"""
        children = {}
        for nm, v in tree.get("items", {}).items():
            if isinstance(v, dict) and ("items" in v or "attrs" in v and "data" not in v):
                children[nm] = self.node(v)[0]
            elif isinstance(v, dict):
                children[nm] = self.dataset(v["data"], v.get("attrs", {}), v.get("chunks"), v.get("gzip"), v.get("shuffle", False))
            else:
                children[nm] = self.dataset(v, {})
        return self.group(children, tree.get("attrs", {}))


def write_h5(path, tree):
    """Write ``tree`` = {'attrs': {name: value}, 'items': {name: ndarray | {'data': ndarray, 'attrs': {...} [, 'chunks':
    (..), 'gzip': level, 'shuffle': bool]} | subtree}} as an HDF5 file (superblock 0, symbol-table groups, contiguous
    or chunked datasets)."""
    w = _Writer()
    w.buf += b"\0" * 96                                    # superblock (56 bytes) + root symbol-table entry (40)
    root, btree, heap = w.node(tree)
    eof = len(w.buf) + (-len(w.buf) % 8)
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, _Writer.LEAF_K, _Writer.NODE_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root, 1, 0) + struct.pack("<QQ", btree, heap)      # cached symbol-table info
    assert len(sb) == 96, len(sb)
    w.buf[:96] = sb
    w.buf += b"\0" * (eof - len(w.buf))
    with open(path, "wb") as fh:
        fh.write(bytes(w.buf))


# ------------------------------------------------------------------------------------------------
# Keras weight files
# ------------------------------------------------------------------------------------------------
def _names(v):
    out = []
    for x in np.asarray(v).ravel():
        out.append(x.decode("utf-8") if isinstance(x, (bytes, np.bytes_)) else str(x))
    return out


def read_keras_weights(path):
    """-> (list of float32 arrays in Keras ``get_weights()`` order, list of 'layer/weight' names).

    Layout written by Keras 2 ``Model.save`` / ``save_weights`` (keras/engine/topology.py, ``save_weights_to_hdf5_group``):
    attribute ``layer_names`` on the weights group (``/model_weights`` in full-model files, the root otherwise), per layer
    a group with attribute ``weight_names`` and the datasets those names point at."""
    f = File(path)
    g = f["model_weights"] if "model_weights" in f else f.root
    if "layer_names" not in g.attrs:
        raise H5Error("%s: no 'layer_names' attribute: not a Keras weight file" % path)
    arrays, names = [], []
    for ln in _names(g.attrs["layer_names"]):
        lg = g[ln]
        wn = lg.attrs.get("weight_names")
        for n in (_names(wn) if wn is not None else []):
            arrays.append(np.asarray(lg[n].read(), dtype=np.float32))
            names.append(ln + "/" + n)
    return arrays, names


def write_keras_weights(path, layers, full_model=True, model_config=None):
    """layers: list of (layer_name, [(weight_name, ndarray), ...]) in model.layers order (layers without weights may be
    listed with an empty list, as Keras does).  full_model: weights under ``/model_weights`` (Model.save) instead of
    the root (save_weights)."""
    items = {}
    for ln, ws in layers:
        sub = {}
        for wn, arr in ws:
            node = sub
            parts = wn.split("/")
            for p in parts[:-1]:
                node = node.setdefault(p, {"items": {}})["items"]
            node[parts[-1]] = np.asarray(arr, dtype=np.float32)
        wn_attr = np.array([wn.encode("utf-8") for wn, _ in ws], dtype="S") if ws else np.zeros((0,), dtype="S1")
        items[ln] = {"attrs": {"weight_names": wn_attr}, "items": sub}
    grp = {"attrs": {"layer_names": np.array([ln.encode("utf-8") for ln, _ in layers], dtype="S"),
                     "backend": b"tensorflow", "keras_version": b"2.1.6"},
           "items": items}
    if full_model:
        attrs = {"keras_version": b"2.1.6", "backend": b"tensorflow"}
        if model_config is not None:
            attrs["model_config"] = model_config
        write_h5(path, {"attrs": attrs, "items": {"model_weights": grp}})
    else:
        write_h5(path, grp)
