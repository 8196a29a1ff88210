"""A small, dependency-free reader / writer for the subset of the HDF5 file format the reference uses on either side of
the hot path: ``Patches.dump`` / ``_load_from_disk`` (/root/reference/tomo_encoders/structures/patches.py:81-110: five
flat datasets written with ``h5py.File.create_dataset(name, data=...)``) and ``DataFile``
(/root/reference/tomo_encoders/structures/datafile.py:196-508: one 3-D dataset, optionally chunked, read and written by
hyperslab).  h5py / libhdf5 are not installable in this image (no network, not in the wheelhouse), so this module
restates the published on-disk format -- "HDF5 File Format Specification Version 2.0", the version-0 superblock layout
every libhdf5 reads and h5py writes by default (libver='earliest'):

    superblock v0 (96 bytes) -> root symbol-table entry -> version-1 object headers
    groups   = symbol-table message (0x0011) -> v1 B-tree (node type 0) -> symbol-table nodes "SNOD" + local heap "HEAP"
    datasets = dataspace (0x0001) + datatype (0x0003) + fill value (0x0005) + data layout v3 (0x0008) messages
    layouts  = contiguous (class 1), chunked (class 2: v1 B-tree, node type 1, every chunk allocated), compact (class 0, read)
    types    = fixed-point (class 0), IEEE floating point incl. float16 (class 1), fixed-length strings (class 3, numpy 'S')

PARITY UNPINNED: no libhdf5 exists here to open these files with, and no HDF5 file to read; the byte layout follows
the specification from memory of it, the tests (tests/test_hdf5_lite.py) check the writer against hand-assembled
known-answer bytes of the superblock / messages and the reader against the writer.  When h5py is importable the
callers (structures/patches.py, structures/datafile.py) use it instead.

The interface mirrors the part of h5py the reference touches: ``File(path, mode)`` as a context manager with
``create_dataset(name, data=, shape=, dtype=, chunks=)``, ``keys()``, ``in``, ``[name]``; datasets with ``shape``,
``dtype``, ``chunks``, numpy-style hyperslab ``[...]`` reads and assignments, ``len`` / iteration over axis 0 and
``np.asarray``.  Reading also follows continuation messages, version-1 superblocks, version-2 dataspaces and unfiltered or
deflate / shuffle-filtered chunks as libhdf5 writes them.
"""
import os
import struct
import zlib

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"
LEAF_K = 4            # symbol-table node holds 2 K entries
INTERNAL_K = 16       # group B-tree node holds 2 K children
CHUNK_K = 32          # chunk B-tree node holds 2 K children (the library's default for superblock v0)
MSG_DATASPACE, MSG_DATATYPE, MSG_FILL_OLD, MSG_FILL, MSG_LAYOUT, MSG_FILTERS = 0x1, 0x3, 0x4, 0x5, 0x8, 0xB
MSG_ATTRIBUTE, MSG_CONTINUATION, MSG_SYMTAB, MSG_MODTIME = 0xC, 0x10, 0x11, 0x12


def _pad8(n):
    return (n + 7) & ~7


# ---------------------------------------------------------------------------------------------- datatypes
def encode_dtype(dt):
    """numpy dtype -> datatype message body (class / version byte, 3 bit-field bytes, size, properties)."""
    dt = np.dtype(dt)
    big = 1 if dt.byteorder == ">" else 0
    if dt.kind in "ui":
        bits0 = big | (0x08 if dt.kind == "i" else 0)
        return struct.pack("<BBBBIHH", 0x10 | 0, bits0, 0, 0, dt.itemsize, 0, 8 * dt.itemsize)
    if dt.kind == "f":
        expsize, mantsize, bias = {2: (5, 10, 15), 4: (8, 23, 127), 8: (11, 52, 1023)}[dt.itemsize]
        nbits = 8 * dt.itemsize
        # bit field: byte order | mantissa normalisation 2 (implied leading 1) in bits 4-5 | sign bit position in bits 8-15
        return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, 0x20 | big, nbits - 1, 0, dt.itemsize, 0, nbits, mantsize, expsize, 0,
                           mantsize, bias)
    if dt.kind == "S":
        # fixed-length string, null-padded (what h5py writes for numpy 'S' arrays), ASCII
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, max(dt.itemsize, 1))
    raise TypeError("hdf5_lite: unsupported dtype %r (integers, float16/32/64 and fixed-length byte strings)" % (dt,))


def decode_dtype(buf, off=0):
    """-> numpy dtype of the datatype message body at buf[off:]."""
    cv, b0, b1, b2, size = struct.unpack_from("<BBBBI", buf, off)
    cls = cv & 0x0F
    order = ">" if (b0 & 1) else "<"
    if cls == 0:
        return np.dtype("%s%s%d" % (order, "i" if (b0 & 0x08) else "u", size))
    if cls == 1:
        if size not in (2, 4, 8):
            raise TypeError("hdf5_lite: floating-point type of %d bytes" % size)
        return np.dtype("%sf%d" % (order, size))
    if cls == 3:
        return np.dtype("S%d" % size)
    raise TypeError("hdf5_lite: datatype class %d is not supported (integers, IEEE floats, fixed-length strings)" % cls)


# ---------------------------------------------------------------------------------------------- object headers
def _message(mtype, body, flags=0):
    body = body + b"\0" * (_pad8(len(body)) - len(body))
    return struct.pack("<HHB3x", mtype, len(body), flags) + body


def _object_header(messages):
    data = b"".join(messages)
    return struct.pack("<BBHII4x", 1, 0, len(messages), 1, len(data)) + data


def _dataspace_body(shape):
    return struct.pack("<BBB5x", 1, len(shape), 0) + b"".join(struct.pack("<Q", int(s)) for s in shape)


FILL_BODY = struct.pack("<BBBBI", 2, 2, 2, 1, 0)    # version 2, late allocation, written if set, default (zero-size) fill value


def dataset_header(shape, dtype, layout):
    """layout: ("contiguous", address, nbytes) or ("chunked", btree_address, chunk_shape)."""
    dtype = np.dtype(dtype)
    if layout[0] == "contiguous":
        lay = struct.pack("<BBQQ", 3, 1, layout[1], layout[2])
    else:
        dims = tuple(int(c) for c in layout[2]) + (dtype.itemsize,)
        lay = struct.pack("<BBBQ", 3, 2, len(dims), layout[1]) + b"".join(struct.pack("<I", d) for d in dims)
    return _object_header([_message(MSG_DATASPACE, _dataspace_body(shape)), _message(MSG_DATATYPE, encode_dtype(dtype), 1),
                           _message(MSG_FILL, FILL_BODY), _message(MSG_LAYOUT, lay)])


def superblock(root_header, root_btree, root_heap, eof, leaf_k=LEAF_K):
    sb = SIGNATURE + struct.pack("<BBBBBBBBHHI", 0, 0, 0, 0, 0, 8, 8, 0, leaf_k, INTERNAL_K, 0)
    sb += struct.pack("<QQQQ", 0, UNDEF, eof, UNDEF)
    sb += struct.pack("<QQII", 0, root_header, 1, 0) + struct.pack("<QQ", root_btree, root_heap)
    assert len(sb) == 96
    return sb


# ---------------------------------------------------------------------------------------------- chunk index
def _chunk_grid(shape, chunks):
    return tuple(-(-int(s) // int(c)) for s, c in zip(shape, chunks))


def _chunk_btree(shape, chunks, itemsize, first_chunk_addr, base):
    """v1 B-tree (node type 1) over densely allocated chunks laid out in C order of their grid index from
    first_chunk_addr.  Returns (bytes, root address); the nodes are placed from `base` upwards."""
    grid = _chunk_grid(shape, chunks)
    nchunks = int(np.prod(grid))
    cbytes = int(np.prod(chunks)) * itemsize
    rank = len(shape)
    offs = np.stack(np.unravel_index(np.arange(nchunks), grid), axis=1) * np.asarray(chunks, dtype=np.int64)
    past = np.zeros(rank, dtype=np.int64)
    past[0] = grid[0] * chunks[0]                                   # a key beyond every chunk of the dataset

    def key(o):
        return struct.pack("<II", cbytes, 0) + b"".join(struct.pack("<Q", int(v)) for v in o) + struct.pack("<Q", 0)

    node_bytes = 24 + (2 * CHUNK_K + 1) * (8 + 8 * (rank + 1)) + 2 * CHUNK_K * 8
    # level 0: children are chunks; entries = (first key offsets, child address)
    entries = [(offs[i], first_chunk_addr + i * cbytes) for i in range(nchunks)]
    out, level, addr = [], 0, base
    while True:
        groups = [entries[i:i + 2 * CHUNK_K] for i in range(0, len(entries), 2 * CHUNK_K)]
        addrs = [addr + j * node_bytes for j in range(len(groups))]
        nxt = []
        for j, g in enumerate(groups):
            left = addrs[j - 1] if j > 0 else UNDEF
            right = addrs[j + 1] if j + 1 < len(groups) else UNDEF
            b = b"TREE" + struct.pack("<BBHQQ", 1, level, len(g), left, right)
            for o, child in g:
                b += key(o) + struct.pack("<Q", child)
            last = groups[j + 1][0][0] if j + 1 < len(groups) else past
            b += key(last)
            out.append(b + b"\0" * (node_bytes - len(b)))
            nxt.append((g[0][0], addrs[j]))
        addr += len(groups) * node_bytes
        if len(groups) == 1:
            return b"".join(out), addrs[0]
        entries, level = nxt, level + 1


# ---------------------------------------------------------------------------------------------- datasets
def _normalise_key(key, shape):
    """-> (tuple of slices / integer-lists per axis, axes to squeeze)."""
    if not isinstance(key, tuple):
        key = (key,)
    if any(k is Ellipsis for k in key):
        i = [j for j, k in enumerate(key) if k is Ellipsis][0]
        key = key[:i] + (slice(None),) * (len(shape) - (len(key) - 1)) + key[i + 1:]
    key = key + (slice(None),) * (len(shape) - len(key))
    if len(key) != len(shape):
        raise IndexError("too many indices for a dataset of rank %d" % len(shape))
    out, squeeze = [], []
    for ax, (k, n) in enumerate(zip(key, shape)):
        if isinstance(k, slice):
            out.append(k)
        elif isinstance(k, (int, np.integer)):
            k = int(k) + (n if k < 0 else 0)
            if not 0 <= k < n:
                raise IndexError("index %d out of range for axis %d of size %d" % (k, ax, n))
            out.append(slice(k, k + 1))
            squeeze.append(ax)
        else:
            out.append(np.asarray(k, dtype=np.int64))            # index list (read_sequence)
    return tuple(out), tuple(squeeze)


class Dataset:
    """One dataset of an open File: numpy-style hyperslab reads and writes straight on the file."""

    def __init__(self, f, name, shape, dtype, layout, filters=()):
        self._f, self.name = f, name
        self.shape, self.dtype = tuple(int(s) for s in shape), np.dtype(dtype)
        self._layout, self._filters = layout, tuple(filters)
        self._index = None

    chunks = property(lambda self: tuple(self._layout[2]) if self._layout[0] == "chunked" else None)
    ndim = property(lambda self: len(self.shape))
    size = property(lambda self: int(np.prod(self.shape)) if self.shape else 1)

    def __len__(self):
        if not self.shape:
            raise TypeError("scalar dataset has no len()")
        return self.shape[0]

    def __iter__(self):
        for i in range(len(self)):
            yield self[i]

    def __array__(self, dtype=None, copy=None):
        a = np.asarray(self[...])
        return a.astype(dtype) if dtype is not None else a

    # -------- contiguous / compact
    def _mapped(self, write):
        kind = self._layout[0]
        if kind == "compact":
            return np.frombuffer(self._layout[1], dtype=self.dtype).reshape(self.shape)
        addr, nbytes = self._layout[1], self._layout[2]
        if addr == UNDEF or self.size == 0:
            return np.zeros(self.shape, self.dtype)                 # storage never allocated: the fill value
        self._f._fh.flush()
        return np.memmap(self._f._fh, dtype=self.dtype, mode="r+" if write else "r", offset=addr, shape=self.shape)

    # -------- chunked
    def _chunk_index(self):
        if self._index is None:
            self._index = {}
            rank = len(self.shape)
            klen = 8 + 8 * (rank + 1)

            def walk(addr):
                hdr = self._f._read(addr, 24)
                if hdr[:4] != b"TREE" or hdr[4] != 1:
                    raise ValueError("hdf5_lite: bad chunk B-tree node at %d" % addr)
                level, used = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
                body = self._f._read(addr + 24, used * (klen + 8) + klen)
                for i in range(used):
                    o = i * (klen + 8)
                    nbytes, mask = struct.unpack_from("<II", body, o)
                    offs = struct.unpack_from("<%dQ" % rank, body, o + 8)
                    child = struct.unpack_from("<Q", body, o + klen)[0]
                    if level == 0:
                        self._index[tuple(int(a) // int(c) for a, c in zip(offs, self._layout[2]))] = (child, nbytes, mask)
                    else:
                        walk(child)
            if self._layout[1] != UNDEF:
                walk(self._layout[1])
        return self._index

    def _read_chunk(self, idx):
        chunks = self._layout[2]
        ent = self._chunk_index().get(idx)
        if ent is None:
            return np.zeros(chunks, self.dtype)
        addr, nbytes, mask = ent
        raw = self._f._read(addr, nbytes)
        for pos in range(len(self._filters) - 1, -1, -1):          # filters are undone in reverse order
            fid = self._filters[pos]
            if mask & (1 << pos):
                continue
            if fid == 1:
                raw = zlib.decompress(raw)
            elif fid == 2:
                n = self.dtype.itemsize
                raw = np.frombuffer(raw, np.uint8).reshape(n, -1).T.tobytes()
            elif fid == 3:
                raw = raw[:-4]
            else:
                raise NotImplementedError("hdf5_lite: filter id %d (deflate 1, shuffle 2, fletcher32 3 are understood)" % fid)
        return np.frombuffer(raw, self.dtype, count=int(np.prod(chunks))).reshape(chunks)

    def _chunked_io(self, sl, value=None):
        chunks = self._layout[2]
        ranges = [range(*s.indices(n)) for s, n in zip(sl, self.shape)]
        out_shape = tuple(len(r) for r in ranges)
        out = None if value is not None else np.empty(out_shape, self.dtype)
        if value is not None:
            if self._filters:
                raise NotImplementedError("hdf5_lite: writing into filtered (compressed) chunks")
            value = np.broadcast_to(np.asarray(value, dtype=self.dtype), out_shape)
        if 0 in out_shape:
            return out
        lo = [r[0] if r.step > 0 else r[-1] for r in ranges]
        hi = [r[-1] if r.step > 0 else r[0] for r in ranges]
        grid_ranges = [range(a // c, b // c + 1) for a, b, c in zip(lo, hi, chunks)]
        for idx in np.ndindex(*[len(g) for g in grid_ranges]):
            cidx = tuple(g[i] for g, i in zip(grid_ranges, idx))
            src, dst, empty = [], [], False
            for ax, (r, c) in enumerate(zip(ranges, chunks)):
                base = cidx[ax] * c
                pos = np.asarray([k for k, v in enumerate(r) if base <= v < base + c], dtype=np.int64) \
                    if (r.step != 1) else None
                if pos is None:
                    a, b = max(r.start, base), min(r.stop, base + c)
                    if a >= b:
                        empty = True
                        break
                    src.append(slice(a - base, b - base))
                    dst.append(slice(a - r.start, b - r.start))
                else:
                    if len(pos) == 0:
                        empty = True
                        break
                    src.append(np.asarray([r[k] - base for k in pos]))
                    dst.append(pos)
            if empty:
                continue
            if value is None:
                ch = self._read_chunk(cidx)
                out[np.ix_(*[np.arange(n)[d] for d, n in zip(dst, out_shape)])] = \
                    ch[np.ix_(*[np.arange(c)[s] for s, c in zip(src, chunks)])]
            else:
                ent = self._chunk_index().get(cidx)
                if ent is None:
                    raise ValueError("hdf5_lite: chunk %s of %r was never allocated" % (cidx, self.name))
                ch = self._read_chunk(cidx).copy()
                ch[np.ix_(*[np.arange(c)[s] for s, c in zip(src, chunks)])] = \
                    value[np.ix_(*[np.arange(n)[d] for d, n in zip(dst, out_shape)])]
                self._f._write(ent[0], ch.tobytes())
        return out

    # -------- numpy-style access
    def _scalar(self):
        if self._layout[0] == "compact":
            return np.frombuffer(self._layout[1], self.dtype, count=1)[0]
        if self._layout[1] == UNDEF:
            return np.zeros((), self.dtype)[()]
        return np.frombuffer(self._f._read(self._layout[1], self.dtype.itemsize), self.dtype)[0]

    def __getitem__(self, key):
        if not self.shape:
            if key not in ((), Ellipsis):
                raise IndexError("scalar dataset indexed with %r" % (key,))
            return self._scalar()
        sl, squeeze = _normalise_key(key, self.shape)
        lists = [ax for ax, s in enumerate(sl) if not isinstance(s, slice)]
        if lists:
            ax = lists[0]
            parts = [self[tuple(slice(int(i), int(i) + 1) if a == ax else s for a, s in enumerate(sl))] for i in sl[ax]]
            out = np.concatenate(parts, axis=ax) if parts else np.empty(tuple(0 if a == ax else len(range(*s.indices(n)))
                                                                                  for a, (s, n) in enumerate(zip(sl, self.shape))), self.dtype)
        elif self._layout[0] == "chunked":
            out = self._chunked_io(sl)
        else:
            out = np.array(self._mapped(False)[sl])
        if squeeze:
            out = out.reshape([n for ax, n in enumerate(out.shape) if ax not in squeeze])
        return out[()] if out.ndim == 0 else out                    # a fully indexed element is a numpy scalar, as in h5py

    def __setitem__(self, key, value):
        if self._f.mode == "r":
            raise OSError("hdf5_lite: file is open read-only")
        if not self.shape:
            self._f._write(self._layout[1], np.asarray(value, dtype=self.dtype).tobytes())
            return
        sl, squeeze = _normalise_key(key, self.shape)
        if any(not isinstance(s, slice) for s in sl):
            raise NotImplementedError("hdf5_lite: assignment through index lists")
        value = np.asarray(value)
        if squeeze and value.ndim == len(self.shape) - len(squeeze):
            value = np.expand_dims(value, squeeze)                   # integer indices dropped these axes on the caller's side
        if self._layout[0] == "chunked":
            self._chunked_io(sl, value)
        elif self._layout[0] == "contiguous":
            if self.size:
                m = self._mapped(True)
                m[sl] = value
                m.flush()
                del m
        else:
            raise NotImplementedError("hdf5_lite: writing into a compact dataset")


class Group(dict):
    """name -> Dataset / Group"""


class File:
    """h5py.File look-alike for the subset above.  mode 'r' | 'r+' | 'w' | 'a'."""

    def __init__(self, path, mode="r"):
        if mode == "a":
            mode = "r+" if os.path.exists(path) else "w"
        if mode not in ("r", "r+", "w"):
            raise ValueError("mode must be 'r', 'r+', 'w' or 'a'")
        self.path, self.mode = path, mode
        self._dirty, self._foreign = False, False
        if mode == "w":
            self._fh = open(path, "w+b")
            self._fh.write(b"\0" * 96)
            self._eof = 96
            self.root = Group()
            self._dirty = True
        else:
            self._fh = open(path, "rb" if mode == "r" else "r+b")
            self._fh.seek(0, os.SEEK_END)
            self._eof = self._fh.tell()
            self.root = self._parse()

    # -------- context manager / mapping interface
    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def close(self):
        if self._fh is None:
            return
        if self._dirty:
            self._write_metadata()
        self._fh.close()
        self._fh = None

    def _walk(self, name, create=False):
        parts = [p for p in str(name).split("/") if p]
        if not parts:
            raise KeyError("empty object name")
        g = self.root
        for p in parts[:-1]:
            if p not in g:
                if not create:
                    raise KeyError(name)
                g[p] = Group()
            g = g[p]
            if not isinstance(g, Group):
                raise KeyError("%r: %r is a dataset, not a group" % (name, p))
        return g, parts[-1]

    def keys(self):
        return self.root.keys()

    def __iter__(self):
        return iter(self.root)

    def __contains__(self, name):
        try:
            g, leaf = self._walk(name)
        except KeyError:
            return False
        return leaf in g

    def __getitem__(self, name):
        g, leaf = self._walk(name)
        if leaf not in g:
            raise KeyError("object %r does not exist" % (name,))
        return g[leaf]

    # -------- raw file access
    def _read(self, addr, n):
        self._fh.seek(addr)
        b = self._fh.read(n)
        if len(b) != n:
            raise ValueError("hdf5_lite: truncated file (%d bytes at %d)" % (n, addr))
        return b

    def _write(self, addr, b):
        self._fh.seek(addr)
        self._fh.write(b)

    def _alloc(self, n):
        addr = _pad8(self._eof)
        self._eof = addr + n
        return addr

    # -------- writing
    def create_dataset(self, name, shape=None, dtype=None, data=None, chunks=None):
        if self.mode == "r":
            raise OSError("hdf5_lite: file is open read-only")
        if self._foreign:
            raise NotImplementedError("hdf5_lite: adding objects to a file that holds features this writer would drop "
                                      "(attributes, filters, ...)")
        g, leaf = self._walk(name, create=True)
        if leaf in g:
            raise ValueError("unable to create dataset %r (name already exists)" % (name,))
        if data is not None:
            data = np.asarray(data)
            if data.dtype.kind == "U":
                raise TypeError("hdf5_lite: numpy unicode arrays have no HDF5 equivalent here; use dtype='S'")
            if dtype is not None:
                data = data.astype(dtype)
            if shape is not None and tuple(shape) != data.shape:
                data = data.reshape(shape)
            shape, dtype = data.shape, data.dtype
        if shape is None or dtype is None:
            raise TypeError("create_dataset needs data or both shape and dtype")
        if isinstance(shape, (int, np.integer)):
            shape = (int(shape),)
        shape, dtype = tuple(int(s) for s in shape), np.dtype(dtype)
        encode_dtype(dtype)                                          # raises for unsupported types before anything is written
        if chunks is True:
            chunks = None
        if chunks is not None:
            chunks = tuple(int(c) for c in chunks)
            if len(chunks) != len(shape) or any(c < 1 for c in chunks):
                raise ValueError("chunk shape %r does not fit dataset shape %r" % (chunks, shape))
            chunks = tuple(min(c, max(s, 1)) for c, s in zip(chunks, shape))
        nbytes = int(np.prod(shape)) * dtype.itemsize if shape else dtype.itemsize
        if chunks is None or nbytes == 0:
            addr = self._alloc(nbytes) if nbytes else UNDEF
            layout = ("contiguous", addr, nbytes)
            if nbytes:
                if data is not None:
                    self._write(addr, np.ascontiguousarray(data).tobytes())
                else:                                                # zero fill by extending the file
                    self._fh.truncate(max(self._eof, addr + nbytes))
        else:
            grid = _chunk_grid(shape, chunks)
            cbytes = int(np.prod(chunks)) * dtype.itemsize
            first = self._alloc(int(np.prod(grid)) * cbytes)
            self._fh.truncate(max(self._eof, first + int(np.prod(grid)) * cbytes))
            tree, root = _chunk_btree(shape, chunks, dtype.itemsize, first, _pad8(self._eof))
            base = self._alloc(len(tree))
            self._write(base, tree)
            layout = ("chunked", root, chunks)
        ds = Dataset(self, leaf, shape, dtype, layout)
        g[leaf] = ds
        self._dirty = True
        if chunks is not None and nbytes and data is not None:
            ds[...] = data
        return ds

    def _write_metadata(self):
        """All object headers, group B-trees, heaps and symbol-table nodes are (re)written as one block at the end of the
        file and the superblock at offset 0 is pointed at the new root group; raw data and chunk indices stay in place."""
        base = _pad8(self._eof)
        blob = bytearray()

        def put(b):
            addr = base + len(blob)
            blob.extend(b)
            blob.extend(b"\0" * (_pad8(len(b)) - len(b)))
            return addr

        max_entries = [1]

        def emit_group(g):
            entries = []
            for name in sorted(g, key=lambda s: s.encode()):
                obj = g[name]
                if isinstance(obj, Group):
                    hdr, bt, hp = emit_group(obj)
                    entries.append((name, hdr, 1, bt, hp))
                else:
                    entries.append((name, put(dataset_header(obj.shape, obj.dtype, obj._layout)), 0, 0, 0))
            max_entries[0] = max(max_entries[0], len(entries))
            heap = bytearray(b"\0" * 8)
            offs = []
            for name, *_ in entries:
                offs.append(len(heap))
                nb = name.encode() + b"\0"
                heap.extend(nb + b"\0" * (_pad8(len(nb)) - len(nb)))
            per = 2 * leaf_k[0]
            snods, keys = [], [0]
            for i in range(0, len(entries), per):
                part = entries[i:i + per]
                b = b"SNOD" + struct.pack("<BBH", 1, 0, len(part))
                for j, (name, hdr, ctype, bt, hp) in enumerate(part):
                    b += struct.pack("<QQII", offs[i + j], hdr, ctype, 0) + struct.pack("<QQ", bt, hp)
                b += b"\0" * (8 + per * 40 - len(b))
                snods.append(put(b))
                keys.append(offs[i + len(part) - 1] if part else 0)
            if len(snods) > 2 * INTERNAL_K:
                raise NotImplementedError("hdf5_lite: more than %d objects in one group" % (2 * INTERNAL_K * per))
            bt = b"TREE" + struct.pack("<BBHQQ", 0, 0, len(snods), UNDEF, UNDEF)
            for k, child in zip(keys[:-1], snods):
                bt += struct.pack("<QQ", k, child)
            bt += struct.pack("<Q", keys[-1])
            bt += b"\0" * (24 + (2 * INTERNAL_K + 1) * 8 + 2 * INTERNAL_K * 8 - len(bt))
            bt_addr = put(bt)
            data_addr = put(bytes(heap))
            hp_addr = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, data_addr))       # free-list head 1 = none
            hdr_addr = put(_object_header([_message(MSG_SYMTAB, struct.pack("<QQ", bt_addr, hp_addr))]))
            return hdr_addr, bt_addr, hp_addr

        def count(g):
            return max([len(g)] + [count(v) for v in g.values() if isinstance(v, Group)])

        leaf_k = [max(LEAF_K, -(-count(self.root) // 2))]
        if leaf_k[0] > 0x7FFF:
            raise NotImplementedError("hdf5_lite: group too large")
        hdr, bt, hp = emit_group(self.root)
        self._write(base, bytes(blob))
        self._eof = base + len(blob)
        self._fh.truncate(self._eof)
        self._write(0, superblock(hdr, bt, hp, self._eof, leaf_k[0]))
        self._fh.flush()
        self._dirty = False

    # -------- reading
    def _messages(self, addr):
        """[(type, flags, body bytes)] of the version-1 object header at addr, continuation blocks included."""
        ver, _, nmsg, _, hsize = struct.unpack_from("<BBHII", self._read(addr, 12))
        if ver != 1:
            raise NotImplementedError("hdf5_lite: object header version %d (files written with libver='latest' use version 2)" % ver)
        out, blocks = [], [(addr + 16, hsize)]
        while blocks and len(out) < nmsg:
            a, n = blocks.pop(0)
            buf, o = self._read(a, n), 0
            while o + 8 <= n and len(out) < nmsg:
                mtype, size, flags = struct.unpack_from("<HHB", buf, o)
                body = buf[o + 8:o + 8 + size]
                o += 8 + size
                if mtype == MSG_CONTINUATION:
                    blocks.append(struct.unpack_from("<QQ", body))
                out.append((mtype, flags, body))
        return out

    def _parse(self):
        head = self._read(0, 8 + 8)
        if head[:8] != SIGNATURE:
            raise OSError("hdf5_lite: %s is not an HDF5 file (signature not at offset 0)" % self.path)
        ver = head[8]
        if ver not in (0, 1):
            raise NotImplementedError("hdf5_lite: superblock version %d (written with libver='latest'); versions 0 and 1 are read" % ver)
        if head[13] != 8 or head[14] != 8:
            raise NotImplementedError("hdf5_lite: %d-byte offsets / %d-byte lengths" % (head[13], head[14]))
        o = 24 + (4 if ver == 1 else 0)
        base, _, eof, _ = struct.unpack_from("<QQQQ", self._read(o, 32))
        if base != 0:
            raise NotImplementedError("hdf5_lite: non-zero base address")
        _, root_hdr, ctype, _ = struct.unpack_from("<QQII", self._read(o + 32, 24))
        self._leaf_k = struct.unpack_from("<H", self._read(16, 2))[0]
        return self._parse_group(root_hdr)

    def _parse_group(self, hdr_addr):
        g = Group()
        msgs = self._messages(hdr_addr)
        sym = [m for m in msgs if m[0] == MSG_SYMTAB]
        if not sym:
            raise NotImplementedError("hdf5_lite: group without a symbol-table message (new-style group of libver='latest')")
        if any(m[0] == MSG_ATTRIBUTE for m in msgs):
            self._foreign = True
        bt_addr, hp_addr = struct.unpack_from("<QQ", sym[0][2])
        hp = self._read(hp_addr, 32)
        if hp[:4] != b"HEAP":
            raise ValueError("hdf5_lite: bad local heap at %d" % hp_addr)
        seg_size, _, seg_addr = struct.unpack_from("<QQQ", hp, 8)
        heap = self._read(seg_addr, seg_size)

        def walk(addr):
            hdr = self._read(addr, 24)
            if hdr[:4] != b"TREE" or hdr[4] != 0:
                raise ValueError("hdf5_lite: bad group B-tree node at %d" % addr)
            level, used = hdr[5], struct.unpack_from("<H", hdr, 6)[0]
            body = self._read(addr + 24, used * 16 + 8)
            for i in range(used):
                child = struct.unpack_from("<Q", body, i * 16 + 8)[0]
                if level > 0:
                    walk(child)
                    continue
                sn = self._read(child, 8)
                if sn[:4] != b"SNOD":
                    raise ValueError("hdf5_lite: bad symbol-table node at %d" % child)
                nsym = struct.unpack_from("<H", sn, 6)[0]
                ents = self._read(child + 8, nsym * 40)
                for j in range(nsym):
                    name_off, obj_hdr, ctype, _ = struct.unpack_from("<QQII", ents, j * 40)
                    name = heap[name_off:heap.index(b"\0", name_off)].decode()
                    g[name] = self._parse_object(name, obj_hdr)
        if bt_addr != UNDEF:
            walk(bt_addr)
        return g

    def _parse_object(self, name, hdr_addr):
        msgs = self._messages(hdr_addr)
        types = {m[0] for m in msgs}
        if MSG_SYMTAB in types:
            return self._parse_group(hdr_addr)
        if MSG_ATTRIBUTE in types or MSG_FILTERS in types:
            self._foreign = True                                     # this writer would drop them when it re-emits the headers
        shape = dtype = layout = None
        filters = []
        for mtype, flags, body in msgs:
            if mtype == MSG_DATASPACE:
                ver, rank, fl = struct.unpack_from("<BBB", body)
                o = 8 if ver == 1 else 4
                shape = struct.unpack_from("<%dQ" % rank, body, o) if rank else ()
            elif mtype == MSG_DATATYPE:
                dtype = decode_dtype(body)
            elif mtype == MSG_FILTERS:
                ver, nf = struct.unpack_from("<BB", body)
                o = 8 if ver == 1 else 2
                for _ in range(nf):
                    fid, = struct.unpack_from("<H", body, o)
                    if ver == 1 or fid >= 256:
                        nlen, fl2, ncv = struct.unpack_from("<HHH", body, o + 2)
                        o += 8 + (_pad8(nlen) if ver == 1 else nlen)
                    else:
                        fl2, ncv = struct.unpack_from("<HH", body, o + 2)
                        o += 6
                    o += 4 * ncv + (4 if (ver == 1 and ncv % 2) else 0)
                    filters.append(fid)
            elif mtype == MSG_LAYOUT:
                ver, cls = struct.unpack_from("<BB", body)
                if ver != 3:
                    raise NotImplementedError("hdf5_lite: data layout message version %d (version 3 is read)" % ver)
                if cls == 1:
                    layout = ("contiguous",) + struct.unpack_from("<QQ", body, 2)
                elif cls == 2:
                    ndim, bt = struct.unpack_from("<BQ", body, 2)
                    dims = struct.unpack_from("<%dI" % ndim, body, 11)
                    layout = ("chunked", bt, tuple(int(d) for d in dims[:-1]))
                elif cls == 0:
                    size, = struct.unpack_from("<H", body, 2)
                    layout = ("compact", bytes(body[4:4 + size]))
                    self._foreign = True
                else:
                    raise NotImplementedError("hdf5_lite: data layout class %d" % cls)
        if shape is None or dtype is None or layout is None:
            raise NotImplementedError("hdf5_lite: object %r is neither a group nor a simple dataset" % name)
        return Dataset(self, name, shape, dtype, layout, filters)


def h5_module():
    """h5py when it is installed (the reference's dependency), else this module: both offer File(path, mode)."""
    try:
        import h5py
        return h5py
    except ImportError:
        import sys
        return sys.modules[__name__]
