"""CPU: the dependency-free HDF5 subset (tomoencoders_b200/structures/hdf5_lite.py), the reference's on-disk schema of
``Patches.dump`` / ``Grid.dump`` (/root/reference/tomo_encoders/structures/patches.py:81-110, grid.py:55-62) and
``DataFile`` (/root/reference/tomo_encoders/structures/datafile.py:83-508) on top of it.

No libhdf5 exists in this image, so interoperability is pinned only as far as bytes can be: the superblock and the
datatype messages are compared with the byte strings libhdf5 / h5py are known to write (the first 96 bytes of every
h5py file with default settings; the IEEE float and integer type descriptions), everything else is the reader against
the writer, plus hand-assembled variants (continuation block, version-1 superblock, filtered chunks) for the reader
paths the writer never produces."""
import os
import struct
import zlib

import numpy as np
import pytest

from tomoencoders_b200 import DataFile, Grid, Patches
from tomoencoders_b200.structures import hdf5_lite as H


def test_superblock_bytes_are_those_of_a_default_h5py_file():
    # root object header at 0x60, its B-tree at 0x88 (0x60 + 40-byte header), the local heap at 0x2a8 (0x88 + 544-byte node):
    # the layout libhdf5 gives every new file
    sb = H.superblock(0x60, 0x88, 0x2A8, 0x1000)
    want = bytes.fromhex("894844460d0a1a0a" "0000000000080800" "04001000" "00000000"
                         "0000000000000000" "ffffffffffffffff" "0010000000000000" "ffffffffffffffff"
                         "0000000000000000" "6000000000000000" "01000000" "00000000"
                         "8800000000000000" "a802000000000000")
    assert sb == want
    # and the writer's own root group has exactly that shape: 16-byte prefix + one 24-byte symbol-table message
    hdr = H._object_header([H._message(H.MSG_SYMTAB, struct.pack("<QQ", 0x88, 0x2A8))])
    assert len(hdr) == 40 and hdr[:16] == bytes.fromhex("01000100" "01000000" "18000000" "00000000")
    assert hdr[16:24] == bytes.fromhex("1100" "1000" "00" "000000")


@pytest.mark.parametrize("dt,hexbytes", [
    ("<f4", "11201f00" "04000000" "0000" "2000" "17" "08" "00" "17" "7f000000"),
    ("<f8", "11203f00" "08000000" "0000" "4000" "34" "0b" "00" "34" "ff030000"),
    ("<f2", "11200f00" "02000000" "0000" "1000" "0a" "05" "00" "0a" "0f000000"),
    ("<i4", "10080000" "04000000" "0000" "2000"),
    ("<u2", "10000000" "02000000" "0000" "1000"),
    ("u1", "10000000" "01000000" "0000" "0800"),
    (">u4", "10010000" "04000000" "0000" "2000"),
    ("S5", "13010000" "05000000"),
])
def test_datatype_messages(dt, hexbytes):
    assert H.encode_dtype(dt) == bytes.fromhex(hexbytes)
    assert H.decode_dtype(bytes.fromhex(hexbytes)) == np.dtype(dt)


def test_dataset_header_layout():
    hdr = H.dataset_header((3, 4), np.uint8, ("contiguous", 0x800, 12))
    ver, _, nmsg, refs, size = struct.unpack_from("<BBHII", hdr)
    assert (ver, nmsg, refs, size) == (1, 4, 1, len(hdr) - 16)
    # dataspace message: type 1, 24 bytes: version 1, rank 2, no max dims, 5 reserved, two 8-byte extents
    assert hdr[16:24] == bytes.fromhex("0100" "1800" "00" "000000")
    assert hdr[24:48] == bytes.fromhex("01" "02" "00" "0000000000") + struct.pack("<QQ", 3, 4)
    # then the datatype (constant flag), the fill value (version 2, late allocation, default value) ...
    assert hdr[48:56] == bytes.fromhex("0300" "1000" "01" "000000") and hdr[56:68] == H.encode_dtype(np.uint8)
    assert hdr[72:80] == bytes.fromhex("0500" "0800" "00" "000000") and hdr[80:88] == bytes.fromhex("02020201" "00000000")
    # ... and the data layout: version 3, contiguous, address, size (18 bytes padded to 24)
    assert hdr[88:96] == bytes.fromhex("0800" "1800" "00" "000000")
    assert hdr[96:120] == bytes.fromhex("0301") + struct.pack("<QQ", 0x800, 12) + b"\0" * 6 and len(hdr) == 120


DTYPES = [np.uint8, np.int8, np.uint16, np.int16, np.uint32, np.int32, np.uint64, np.int64, np.float16, np.float32, np.float64]


def test_round_trip_of_every_supported_type_scalars_strings_and_groups(tmp_path):
    f = str(tmp_path / "a.h5")
    rng = np.random.default_rng(0)
    arrays = {"d_%s" % np.dtype(dt).name: (rng.uniform(-100, 100, (5, 7, 3)) if np.dtype(dt).kind != "u" else rng.uniform(0, 200, (5, 7, 3))).astype(dt)
              for dt in DTYPES}
    with H.File(f, "w") as hf:
        for k, v in arrays.items():
            hf.create_dataset(k, data=v)
        hf.create_dataset("scalar", data=64)
        hf.create_dataset("names", data=np.asarray(["alpha", "b", ""], dtype="S"))
        hf.create_dataset("empty", data=np.zeros((0, 3), np.float32))
        hf.create_dataset("exchange/data", data=arrays["d_uint16"])
        hf.create_dataset("exchange/deeper/still", data=np.arange(4))
        with pytest.raises(ValueError):
            hf.create_dataset("scalar", data=1)
        with pytest.raises(TypeError):
            hf.create_dataset("u", data=np.asarray(["x"]))
    with H.File(f, "r") as hf:
        assert sorted(hf.keys()) == sorted(list(arrays) + ["scalar", "names", "empty", "exchange"])     # 15 objects > 2 * LEAF_K
        for k, v in arrays.items():
            assert hf[k].dtype == v.dtype and hf[k].shape == v.shape and hf[k].chunks is None
            np.testing.assert_array_equal(hf[k][...], v)
            np.testing.assert_array_equal(np.asarray(hf[k]), v)
            np.testing.assert_array_equal(hf[k][1:4, ::2, -1], v[1:4, ::2, -1])
            np.testing.assert_array_equal(hf[k][[0, 3, 4], ...], v[[0, 3, 4]])
        assert int(hf["scalar"][()]) == 64 and hf["scalar"].shape == () and int(np.asarray(hf["scalar"])) == 64
        assert [s.decode("UTF-8") for s in hf["names"]] == ["alpha", "b", ""]
        assert hf["empty"].shape == (0, 3) and hf["empty"][...].shape == (0, 3)
        assert "exchange/data" in hf and "exchange/nothing" not in hf and "nothing" not in hf
        np.testing.assert_array_equal(hf["exchange/data"][...], arrays["d_uint16"])
        np.testing.assert_array_equal(hf["exchange"]["deeper"]["still"][...], np.arange(4))
        with pytest.raises(KeyError):
            hf["nothing"]
        with pytest.raises(OSError):
            hf["scalar"][()] = 3
    # the leaf K of the superblock grew with the widest group so that libhdf5 would find all its entries in one node
    with open(f, "rb") as fh:
        head = fh.read(24)
    assert struct.unpack_from("<H", head, 16)[0] >= 8
    assert os.path.getsize(f) == struct.unpack_from("<Q", open(f, "rb").read(48), 40)[0]        # end-of-file address


@pytest.mark.parametrize("shape,chunks", [((20, 17, 33), (8, 8, 8)), ((70, 9, 9), (1, 9, 9)), ((5, 6), (5, 6)), ((130,), (2,))])
def test_chunked_datasets_with_edge_chunks_and_multi_level_index(tmp_path, shape, chunks):
    f = str(tmp_path / "c.h5")
    ref = np.arange(int(np.prod(shape)), dtype=np.uint16).reshape(shape)
    with H.File(f, "w") as hf:
        d = hf.create_dataset("vol", shape=shape, dtype=np.uint16, chunks=chunks)
        assert d.chunks == chunks
        np.testing.assert_array_equal(d[...], np.zeros(shape, np.uint16))       # allocated and zero-filled
        d[...] = ref
    with H.File(f, "r+") as hf:
        d = hf["vol"]
        assert d.chunks == chunks and d.shape == shape
        np.testing.assert_array_equal(d[...], ref)
        key = (slice(3, None, 2),) + (slice(1, -1),) * (len(shape) - 1)
        np.testing.assert_array_equal(d[key], ref[key])
        np.testing.assert_array_equal(d[-1], ref[-1])
        d[2:4] = 7
        ref[2:4] = 7
        if len(shape) == 3:
            d[:, 3, :] = ref[:, 3, :] + 1
            ref[:, 3, :] += 1
    with H.File(f, "r") as hf:
        np.testing.assert_array_equal(hf["vol"][...], ref)


def test_appending_a_dataset_and_writing_in_place(tmp_path):
    f = str(tmp_path / "b.h5")
    a = np.arange(24, dtype=np.float32).reshape(2, 3, 4)
    with H.File(f, "w") as hf:
        hf.create_dataset("a", data=a)
    with H.File(f, "a") as hf:
        hf.create_dataset("b", shape=(3, 3), dtype=np.uint8)
        hf["b"][1] = [1, 2, 3]
        hf["a"][:, 0, :] = -1.0
    a[:, 0, :] = -1.0
    with H.File(f, "r") as hf:
        np.testing.assert_array_equal(hf["a"][...], a)
        np.testing.assert_array_equal(hf["b"][...], [[0, 0, 0], [1, 2, 3], [0, 0, 0]])
    open(str(tmp_path / "x.bin"), "wb").write(b"abc" * 40)
    with pytest.raises(OSError):
        H.File(str(tmp_path / "x.bin"))


def _hand_built(tmp_path, variant):
    """A file assembled by hand around one dataset, for reader paths the writer never produces."""
    data = np.arange(64, dtype="<i4").reshape(8, 8)
    blob = bytearray(b"\0" * (100 if variant == "sb1" else 96))

    def put(b):
        while len(blob) % 8:
            blob.append(0)
        a = len(blob)
        blob.extend(b)
        return a
    if variant == "filtered":
        raw = data.tobytes()
        shuffled = np.frombuffer(raw, np.uint8).reshape(-1, 4).T.tobytes()           # shuffle, then deflate
        comp = zlib.compress(shuffled)
        caddr = put(comp)
        key = lambda nb, o: struct.pack("<II", nb, 0) + struct.pack("<QQQ", o[0], o[1], 0)    # noqa: E731
        node = b"TREE" + struct.pack("<BBHQQ", 1, 0, 1, H.UNDEF, H.UNDEF) + key(len(comp), (0, 0)) + struct.pack("<Q", caddr) + key(0, (8, 0))
        bt = put(node)
        lay = struct.pack("<BBBQ", 3, 2, 3, bt) + struct.pack("<III", 8, 8, 4)
        # filter pipeline version 1: shuffle (id 2, one client value) then deflate (id 1, one client value)
        filt = struct.pack("<BB6x", 1, 2) + struct.pack("<HHHH", 2, 0, 0, 1) + struct.pack("<I4x", 4) \
            + struct.pack("<HHHH", 1, 0, 1, 1) + struct.pack("<I4x", 6)
        msgs = [H._message(H.MSG_DATASPACE, H._dataspace_body((8, 8))), H._message(H.MSG_DATATYPE, H.encode_dtype("<i4"), 1),
                H._message(H.MSG_FILTERS, filt), H._message(H.MSG_LAYOUT, lay)]
        ds = put(H._object_header(msgs))
    else:
        daddr = put(data.tobytes())
        lay = struct.pack("<BBQQ", 3, 1, daddr, data.nbytes)
        if variant == "continuation":
            # version-2 dataspace in the second block, reached through a continuation message
            second = H._message(H.MSG_LAYOUT, lay) + H._message(H.MSG_DATASPACE, struct.pack("<BBBB", 2, 2, 0, 1) + struct.pack("<QQ", 8, 8))
            caddr = put(second)
            first = [H._message(H.MSG_DATATYPE, H.encode_dtype("<i4"), 1), H._message(0, b"\0" * 8),
                     H._message(H.MSG_CONTINUATION, struct.pack("<QQ", caddr, len(second)))]
            hdr = b"".join(first)
            ds = put(struct.pack("<BBHII4x", 1, 0, 5, 1, len(hdr)) + hdr)
        else:
            ds = put(H.dataset_header((8, 8), "<i4", ("contiguous", daddr, data.nbytes)))
    heap = b"\0" * 8 + b"data\0\0\0\0"
    snod = put(b"SNOD" + struct.pack("<BBH", 1, 0, 1) + struct.pack("<QQII16x", 8, ds, 0, 0) + b"\0" * (7 * 40))
    bt = put(b"TREE" + struct.pack("<BBHQQ", 0, 0, 1, H.UNDEF, H.UNDEF) + struct.pack("<QQQ", 0, snod, 8) + b"\0" * (544 - 48))
    hd = put(heap)
    hp = put(b"HEAP" + struct.pack("<B3xQQQ", 0, len(heap), 1, hd))
    root = put(H._object_header([H._message(H.MSG_SYMTAB, struct.pack("<QQ", bt, hp))]))
    sb = H.superblock(root, bt, hp, len(blob))
    if variant == "sb1":        # version 1: "indexed storage internal node K" + 2 reserved bytes after the consistency flags
        sb = sb[:8] + b"\x01" + sb[9:24] + struct.pack("<HH", 32, 0) + sb[24:]
    blob[:len(sb)] = sb
    path = str(tmp_path / (variant + ".h5"))
    open(path, "wb").write(bytes(blob))
    return path, data


@pytest.mark.parametrize("variant", ["plain", "continuation", "sb1", "filtered"])
def test_reader_on_hand_assembled_files(tmp_path, variant):
    path, data = _hand_built(tmp_path, variant)
    with H.File(path, "r") as hf:
        assert list(hf.keys()) == ["data"]
        d = hf["data"]
        assert d.shape == (8, 8) and d.dtype == np.dtype("<i4")
        np.testing.assert_array_equal(d[...], data)
        np.testing.assert_array_equal(d[2:5, ::3], data[2:5, ::3])


def test_patches_and_grid_dump_load_in_the_reference_schema(tmp_path):
    g = Grid((128, 256, 256), width=64)
    f = str(tmp_path / "g.hdf5")
    g.dump(f)
    with H.File(f, "r") as hf:                                  # grid.py:55-62: vol_shape, points, width
        assert sorted(hf.keys()) == ["points", "vol_shape", "width"] and hf["width"].shape == ()
    g2 = Grid((1, 1, 1), "file", fpath=f)
    np.testing.assert_array_equal(g2.points, g.points)
    assert g2.points.dtype == np.uint32 and g2.vol_shape == g.vol_shape and g2.wd == 64
    p = Patches((128, 256, 256), "regular-grid", patch_size=(64,) * 3)
    p.add_features(np.stack([np.arange(len(p)), np.arange(len(p)) ** 2], axis=1).astype(np.float32), names=["a", "bb"])
    f = str(tmp_path / "p.hdf5")
    p.dump(f)
    with H.File(f, "r") as hf:                                  # patches.py:81-91
        assert sorted(hf.keys()) == ["feature_names", "features", "points", "vol_shape", "widths"]
        assert hf["feature_names"].dtype == np.dtype("S2") and hf["points"].dtype == np.uint32
    p2 = Patches((1, 1, 1), "file", fpath=f)
    np.testing.assert_array_equal(p2.points, p.points)
    np.testing.assert_array_equal(p2.widths, p.widths)
    np.testing.assert_array_equal(p2.features, p.features)
    assert p2.feature_names == ["a", "bb"] and p2.vol_shape == p.vol_shape
    q = Patches((64, 64, 64), "regular-grid", patch_size=(32,) * 3)          # no features: the two optional datasets are absent
    q.dump(str(tmp_path / "q.h5"))
    q2 = Patches((1, 1, 1), "file", fpath=str(tmp_path / "q.h5"))
    assert q2.features is None and q2.feature_names == []


@pytest.mark.parametrize("chunk_shape", [None, (8, 16, 16)])
def test_datafile_hdf5(tmp_path, chunk_shape):
    f = str(tmp_path / "vol.hdf5")
    vol = np.random.default_rng(1).integers(0, 60000, (24, 40, 48)).astype(np.uint16)
    with pytest.raises(ValueError):
        DataFile(f, data_tag="exchange/data", VERBOSITY=0)                   # new file needs d_shape and d_type
    df = DataFile(f, data_tag="exchange/data", d_shape=vol.shape, d_type=vol.dtype, chunk_shape=chunk_shape, VERBOSITY=0)
    assert df.chunk_shape == chunk_shape and abs(df.d_size_GB - vol.nbytes / 1e9) < 1e-12
    df.create_new()
    with pytest.raises(ValueError):
        df.create_new(overwrite=False)
    df.write_chunk(vol[:10], axis=0, s=slice(0, 10))
    df.write_chunk(vol[10:], axis=0, s=slice(10, 24))
    d2 = DataFile(f, data_tag="exchange/data", VERBOSITY=0)                   # existing file: stats are read back
    assert d2.d_shape == vol.shape and d2.d_type == vol.dtype and d2.chunk_shape == chunk_shape
    np.testing.assert_array_equal(d2.read_full(), vol)
    np.testing.assert_array_equal(d2.read_slice(axis=1, slice_idx=7), vol[:, 7, :])
    np.testing.assert_array_equal(d2.read_slice(axis=2, slice_idx=47), vol[:, :, 47])
    ch, s = d2.read_chunk(axis=2, slice_start=4, slice_end=20, skip_fac=3)
    np.testing.assert_array_equal(ch, np.moveaxis(vol[:, :, 4:20:3], 2, 0))
    assert s == slice(4, 20, 3)
    ch, s = d2.read_chunk(axis=0, slice_start=5, max_GB=3 * d2.slice_size[0])
    clen = chunk_shape[0] if chunk_shape else 1
    want = max(1, 3 // clen) * clen
    assert ch.shape[0] == min(want, 24 - 5)
    np.testing.assert_array_equal(d2.read_data((slice(2, 9), slice(None), slice(10, 20))), vol[2:9, :, 10:20])
    np.testing.assert_array_equal(d2.read_sequence([1, 4, 9]), vol[[1, 4, 9]])
    np.testing.assert_array_equal(d2.read_full(skip_fac=5), vol[::5])
    # axis-1 and axis-2 chunks, a block write
    d2.write_chunk(np.moveaxis(vol[:, 3:6, :] + 1, 1, 0), axis=1, s=slice(3, 6))
    d2.write_data(np.zeros((2, 2, 2), np.uint16), slice_3D=(slice(0, 2), slice(0, 2), slice(0, 2)))
    vol2 = vol.copy()
    vol2[:, 3:6, :] += 1
    vol2[:2, :2, :2] = 0
    np.testing.assert_array_equal(d2.read_full(), vol2)
    df.create_new(overwrite=True)
    assert DataFile(f, data_tag="exchange/data", VERBOSITY=0).read_full().max() == 0
    assert "Dataset shape" in d2.show_stats(return_output=True) and "Found existing hdf5 file" in d2.get_stats(return_output=True)


def test_datafile_chunk_estimates_follow_the_reference_rules(tmp_path):
    kw = dict(data_tag="d", d_shape=(100, 200, 300), d_type=np.float32, VERBOSITY=0)
    size = 100 * 200 * 300 * 4 / 1e9
    df = DataFile(str(tmp_path / "x.h5"), chunk_size=size / 8, **kw)
    assert df.chunk_shape == (50, 100, 150)                                   # cube-root scaling of every axis
    df = DataFile(str(tmp_path / "x.h5"), chunked_slice_size=size / 10, **kw)
    assert df.chunk_shape == (10, 20, 30)                                     # that many slices along each axis
    with pytest.raises(ValueError):
        DataFile(str(tmp_path / "x.h5"), chunked_slice_size=2 * size, **kw)
    with pytest.raises(ValueError):
        DataFile(str(tmp_path / "x.h5"), data_tag="d", d_shape=(4, 4, 4), d_type=np.float64, VERBOSITY=0)
    assert DataFile(str(tmp_path / "x.h5"), **kw).chunk_shape is None


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.float32])
def test_datafile_tiff_series(tmp_path, dt):
    folder = str(tmp_path / "rec")
    vol = (np.random.default_rng(2).uniform(0, 250, (12, 20, 24))).astype(dt)
    df = DataFile(folder, tiff=True, d_shape=vol.shape, d_type=vol.dtype, VERBOSITY=0)
    df.write_chunk(vol[:5], axis=0, s=slice(0, 5))
    df.write_chunk(vol[5:], axis=0, s=slice(5, 12))
    names = sorted(os.listdir(folder))
    assert names[0] == "rec01.tif" and names[-1] == "rec12.tif" and len(names) == 12       # datafile.py:578-581 naming
    with pytest.raises(FileExistsError):
        df.write_chunk(vol[:5], axis=0, s=slice(0, 5))
    with pytest.raises(ValueError):
        df.write_chunk(vol[:5], axis=1, s=slice(0, 5))
    d2 = DataFile(folder, tiff=True, VERBOSITY=0)
    assert d2.d_shape == vol.shape and d2.d_type == vol.dtype and d2.chunk_shape is None
    np.testing.assert_array_equal(d2.read_full(), vol)
    np.testing.assert_array_equal(d2.read_slice(axis=0, slice_idx=3), vol[3])
    np.testing.assert_array_equal(d2.read_data((slice(2, 8), slice(1, 5), slice(None))), vol[2:8, 1:5, :])
    np.testing.assert_array_equal(d2.read_sequence([0, 11]), vol[[0, 11]])
    auto = DataFile(folder, autodetect_format=True, VERBOSITY=0)
    assert auto.tiff_mode and auto.d_shape == vol.shape
    d2.create_new(overwrite=True)
    assert not os.path.exists(folder)


def test_random_hyperslabs_match_numpy(tmp_path):
    """Property check (seeded): reads and writes through arbitrary slices -- positive steps, negative bounds, integers,
    partially out-of-range stops -- of contiguous and chunked datasets behave like the same operations on a numpy array."""
    rng = np.random.default_rng(123)

    def rand_key(shape):
        key = []
        for n in shape:
            kind = rng.integers(0, 4)
            if kind == 0:
                key.append(slice(None))
            elif kind == 1:
                key.append(int(rng.integers(-n, n)))
            else:
                a, b = sorted(int(v) for v in rng.integers(-n - 2, n + 3, 2))
                key.append(slice(a, b, int(rng.integers(1, 4)) if kind == 3 else None))
        return tuple(key)

    for trial in range(12):
        rank = int(rng.integers(1, 4))
        shape = tuple(int(v) for v in rng.integers(1, 14, rank))
        chunks = None if trial % 2 == 0 else tuple(int(rng.integers(1, s + 1)) for s in shape)
        ref = rng.integers(0, 1000, shape).astype(np.int32)
        f = str(tmp_path / ("p%d.h5" % trial))
        with H.File(f, "w") as hf:
            hf.create_dataset("d", data=ref, chunks=chunks)
        with H.File(f, "r+") as hf:
            d = hf["d"]
            for _ in range(25):
                key = rand_key(shape)
                got, want = d[key], ref[key]
                assert np.shape(got) == np.shape(want), (shape, chunks, key)
                np.testing.assert_array_equal(got, want)
                if rng.integers(0, 3) == 0 and np.size(want):
                    val = rng.integers(0, 1000, np.shape(want)).astype(np.int32)
                    d[key] = val
                    ref[key] = val
        with H.File(f, "r") as hf:
            np.testing.assert_array_equal(hf["d"][...], ref)


def test_written_file_walks_like_the_specification_says(tmp_path):
    """An independent structural walk of a written file (no code shared with the module's reader): every address the
    superblock, the symbol-table entries, the B-tree and the heap hold must lead to the object the specification puts there."""
    f = str(tmp_path / "s.h5")
    a = np.arange(30, dtype=np.int16).reshape(5, 6)
    with H.File(f, "w") as hf:
        hf.create_dataset("zeta", data=a)
        hf.create_dataset("alpha", data=np.float64(2.5))
        hf.create_dataset("mid/inner", data=np.arange(3, dtype=np.uint8))
    raw = open(f, "rb").read()
    u16 = lambda o: struct.unpack_from("<H", raw, o)[0]        # noqa: E731
    u64 = lambda o: struct.unpack_from("<Q", raw, o)[0]        # noqa: E731
    assert raw[:8] == b"\x89HDF\r\n\x1a\n" and raw[8] == 0 and raw[13] == 8 and raw[14] == 8
    assert u64(24) == 0 and u64(32) == 2 ** 64 - 1 and u64(40) == len(raw) and u64(48) == 2 ** 64 - 1
    root_hdr, cache_type, bt, hp = u64(64), struct.unpack_from("<I", raw, 72)[0], u64(80), u64(88)
    assert u64(56) == 0 and cache_type == 1 and root_hdr % 8 == 0

    def group(hdr, bt_expected, hp_expected):
        # object header: version 1, one message, the symbol-table message pointing at the same B-tree / heap as the entry
        assert raw[hdr] == 1 and u16(hdr + 2) == 1 and struct.unpack_from("<I", raw, hdr + 8)[0] == 24
        assert u16(hdr + 16) == 0x11 and u16(hdr + 18) == 16 and (u64(hdr + 24), u64(hdr + 32)) == (bt_expected, hp_expected)
        assert raw[hp_expected:hp_expected + 4] == b"HEAP" and raw[hp_expected + 4] == 0
        seg_size, free_head, seg = u64(hp_expected + 8), u64(hp_expected + 16), u64(hp_expected + 24)
        assert free_head == 1 and seg_size % 8 == 0 and raw[seg:seg + 8] == b"\0" * 8
        name = lambda off: raw[seg + off:raw.index(b"\0", seg + off)].decode()      # noqa: E731
        assert raw[bt_expected:bt_expected + 4] == b"TREE" and raw[bt_expected + 4] == 0 and raw[bt_expected + 5] == 0
        used = u16(bt_expected + 6)
        assert u64(bt_expected + 8) == u64(bt_expected + 16) == 2 ** 64 - 1 and u64(bt_expected + 24) == 0
        out = {}
        for i in range(used):
            snod = u64(bt_expected + 24 + 16 * i + 8)
            assert raw[snod:snod + 4] == b"SNOD" and raw[snod + 4] == 1
            n = u16(snod + 6)
            names = []
            for j in range(n):
                e = snod + 8 + 40 * j
                names.append(name(u64(e)))
                out[names[-1]] = (u64(e + 8), struct.unpack_from("<I", raw, e + 16)[0], u64(e + 24), u64(e + 32))
            assert names == sorted(names)
            assert name(u64(bt_expected + 24 + 16 * (i + 1))) == names[-1]       # right key = largest name of the child
        return out

    root = group(root_hdr, bt, hp)
    assert sorted(root) == ["alpha", "mid", "zeta"]
    assert root["mid"][1] == 1 and sorted(group(root["mid"][0], root["mid"][2], root["mid"][3])) == ["inner"]
    # the dataset "zeta": four messages; the contiguous layout's address holds the array bytes
    hdr = root["zeta"][0]
    assert raw[hdr] == 1 and u16(hdr + 2) == 4 and root["zeta"][1] == 0
    o, kinds = hdr + 16, []
    for _ in range(4):
        mtype, size = u16(o), u16(o + 2)
        assert size % 8 == 0
        kinds.append(mtype)
        if mtype == 8:
            assert raw[o + 8] == 3 and raw[o + 9] == 1
            addr, nbytes = u64(o + 10), u64(o + 18)
            assert nbytes == a.nbytes and raw[addr:addr + nbytes] == a.tobytes()
        o += 8 + size
    assert kinds == [1, 3, 5, 8] and o - (hdr + 16) == struct.unpack_from("<I", raw, hdr + 8)[0]
