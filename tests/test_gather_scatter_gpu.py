"""GPU: gather / scatter / min-max kernels through the reference-facing API, bit-exact against the
golden vectors, the numpy oracle and size-independent properties."""
import numpy as np
import pytest
import torch

from conftest import seeded_volume
from oracle import structures_np as S, voxel_np as V
from tomoencoders_b200 import Grid, Patches, voxel_processing

pytestmark = pytest.mark.gpu
SVS = (40, 44, 48)
DTYPES = (("u8", np.uint8), ("u16", np.uint16), ("f16", np.float16), ("f32", np.float32))


@pytest.fixture(autouse=True, params=["default", "tma", "simt", "chunks", "simt-rows"])
def gather_path(request, monkeypatch):
    """Every test runs five times: default kernel selection (TMA-staged kernels for 16-byte-aligned
    2- and 4-byte tilings and for short rows at random corners of a volume whose row pitch is a multiple of
    16 bytes, SIMT otherwise), TMA kernels forced wherever the row pitch allows (covers
    their per-unit unaligned / binned paths), SIMT kernels forced, the row kernels switched off, and the
    TMA row kernel switched off (the SIMT row-group kernel takes its cases).  The variables are read on
    every call."""
    monkeypatch.setenv("TOMOENC_NO_TMA_GATHER", "1" if request.param == "simt" else "0")
    monkeypatch.setenv("TOMOENC_FORCE_TMA_GATHER", "1" if request.param == "tma" else "0")
    # "chunks": the row-group kernel for short unaligned rows switched off (the 16-byte chunk kernel takes them)
    monkeypatch.setenv("TOMOENC_NO_ROW_GATHER", "1" if request.param == "chunks" else "0")
    monkeypatch.setenv("TOMOENC_NO_TMA_ROW_GATHER", "1" if request.param == "simt-rows" else "0")
    return request.param


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view({1: np.uint8, 2: np.uint16, 4: np.uint32}[a.dtype.itemsize])


@pytest.mark.parametrize("name,dt", DTYPES)
def test_golden_vectors(golden, name, dt):
    vol = seeded_volume(SVS, dt, 100)
    pg = Patches(SVS, "grid", patch_size=(16,) * 3, stride=1)
    np.testing.assert_array_equal(pg.points, golden["gs_%s_grid_points" % name])
    out = pg.extract(vol, (16,) * 3)
    assert isinstance(out, np.ndarray) and out.dtype == vol.dtype and out.shape == (len(pg), 16, 16, 16)
    np.testing.assert_array_equal(_bits(out), _bits(golden["gs_%s_grid_extract" % name]))
    pb = Patches(SVS, "data", points=golden["gs_%s_bin_points" % name], widths=golden["gs_%s_bin_widths" % name])
    np.testing.assert_array_equal(_bits(pb.extract(vol, (16,) * 3)), _bits(golden["gs_%s_bin_extract" % name]))
    sub = seeded_volume((len(pg),) + (16,) * 3, dt, 101)
    vout = seeded_volume(SVS, dt, 102)
    pg.fill_patches_in_volume(sub, vout)          # overlapping patches: last index wins
    np.testing.assert_array_equal(_bits(vout), _bits(golden["gs_%s_fill" % name]))


def test_grid_extract_golden_and_round_trip(golden):
    gv = seeded_volume((32, 48, 32), np.uint8, 103)
    g = Grid(gv.shape, width=16)
    x = g.extract(gv)
    np.testing.assert_array_equal(x, golden["grid_extract_u8"])
    out = np.zeros_like(gv)
    g.fill_patches_in_volume(x, out)
    np.testing.assert_array_equal(out, gv)


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.float16, np.float32])
def test_aligned_pitch_random_corners_and_mixed_binning(dt):
    """Row pitch a multiple of 16 bytes -> TMA box loads at arbitrary (odd) corners; the 'random'
    initialiser mixes binning factors 1..3 in one call (binned units take the in-kernel slow path)."""
    vs = (96, 100, 128)
    vol = seeded_volume(vs, dt, 12)
    np.random.seed(5)
    p = Patches(vs, "random-fixed-width", patch_size=(32,) * 3, n_points=301)
    assert np.any(p.points % 2 == 1)
    np.testing.assert_array_equal(_bits(p.extract(vol, (32,) * 3)), _bits(S.extract(vol, p.points, p.widths, (32,) * 3)))
    np.random.seed(6)
    q = Patches(vs, "random", min_patch_size=(16,) * 3, max_stride=4, n_points=97)
    assert len(np.unique(q.widths[:, 0])) > 1
    np.testing.assert_array_equal(_bits(q.extract(vol, (16,) * 3)), _bits(S.extract(vol, q.points, q.widths, (16,) * 3)))
    # non-cubic patches, several z chunks per patch
    r = Patches(vs, "data", points=[[1, 2, 3], [40, 30, 61], [0, 0, 0]], widths=[[48, 40, 64]] * 3)
    np.testing.assert_array_equal(_bits(r.extract(vol, (48, 40, 64))), _bits(S.extract(vol, r.points, r.widths, (48, 40, 64))))
    # disjoint scatter at unaligned corners (explicit points, so the lattice test of the host code
    # fails and the ordered path is used) and on the lattice (TMA box stores)
    g = Grid(vs[:1] + (96, 128), width=32)
    sub = seeded_volume((len(g), 32, 32, 32), dt, 13)
    out = seeded_volume(vs[:1] + (96, 128), dt, 14)
    exp = out.copy()
    g.fill_patches_in_volume(sub, out)
    S.fill_patches_in_volume(sub, g.points, np.full((len(g), 3), 32), exp)
    np.testing.assert_array_equal(_bits(out), _bits(exp))


@pytest.mark.parametrize("dt", [np.uint8, np.uint16, np.float32])
def test_row_gather_edges_and_odd_row_lengths(dt):
    """Random-corner row kernels (TMA-staged box 16 bytes wider than the row, start rounded down): patches that touch
    the far faces of the volume (the padded box hangs over the edge: out-of-bounds fill), every byte phase of the
    corner, rows of 3 / 7 / 14 chunks (the generic chunks-per-row path), several units per patch and one unit for
    several patches' worth of planes."""
    vs = (70, 66, 160)
    vol = seeded_volume(vs, dt, 31)
    es = np.dtype(dt).itemsize
    for patch in ((8, 8, 48 // es), (5, 9, 112 // es), (64, 64, 32), (3, 2, 16)):
        pz, py, px = patch
        if px * es % 16 or px * es > 240:
            continue
        g = np.random.default_rng(pz)
        pts = np.stack([g.integers(0, vs[a] - patch[a] + 1, 200) for a in range(3)], axis=1)
        pts[:40, 2] = vs[2] - px                # right face
        pts[40:60, 0] = vs[0] - pz              # last planes
        pts[60:80, 1] = vs[1] - py
        pts[80:96, 2] = np.arange(16)           # every byte phase
        pts[96] = (vs[0] - pz, vs[1] - py, vs[2] - px)
        pts[97] = 0
        p = Patches(vs, "data", points=pts, widths=np.tile(np.array(patch), (len(pts), 1)))
        np.testing.assert_array_equal(_bits(p.extract(vol, patch)), _bits(S.extract(vol, p.points, p.widths, patch)))
        tv = torch.from_numpy(_bits(vol).astype({1: np.uint8, 2: np.int16, 4: np.int32}[es], copy=False)).cuda()
        out = p.extract(tv, patch)             # device in, device out
        assert out.is_cuda
        np.testing.assert_array_equal(_bits(out.cpu().numpy()), _bits(S.extract(vol, p.points, p.widths, patch)))


@pytest.mark.parametrize("dt", [np.uint8, np.float16, np.float32])
def test_random_unaligned_corners_match_oracle(dt):
    vs = (96, 100, 117)                       # odd pitch: rows start at every byte alignment
    vol = seeded_volume(vs, dt, 11)
    np.random.seed(3)
    p = Patches(vs, "random-fixed-width", patch_size=(32,) * 3, n_points=257)
    out = p.extract(vol, (32,) * 3)
    np.testing.assert_array_equal(_bits(out), _bits(S.extract(vol, p.points, p.widths, (32,) * 3)))
    np.random.seed(4)
    q = Patches(vs, "random", min_patch_size=(16,) * 3, max_stride=4, n_points=64)
    np.testing.assert_array_equal(_bits(q.extract(vol, (16,) * 3)), _bits(S.extract(vol, q.points, q.widths, (16,) * 3)))
    # rows that are not a multiple of 16 bytes take the element-wise kernel
    r = Patches(vs, "data", points=[[1, 2, 3], [60, 70, 80]], widths=[[20, 24, 28]] * 2)
    np.testing.assert_array_equal(_bits(r.extract(vol, (20, 24, 28))), _bits(S.extract(vol, r.points, r.widths, (20, 24, 28))))


def test_device_in_device_out_and_list_mode():
    vol = seeded_volume((64, 64, 64), np.uint8, 12)
    tv = torch.from_numpy(vol).cuda()
    p = Patches(vol.shape, "data", points=[[0, 0, 0], [10, 20, 30], [5, 5, 5]], widths=[[32] * 3, [16] * 3, [32] * 3])
    lst = p.extract(tv, None)
    assert isinstance(lst, list) and all(t.is_cuda for t in lst)
    for i, t in enumerate(lst):
        z, y, x = p.points[i].astype(int)
        w = int(p.widths[i, 0])
        np.testing.assert_array_equal(t.cpu().numpy(), vol[z:z + w, y:y + w, x:x + w])
    g = Grid(vol.shape, width=32)
    x = g.extract(tv)
    assert x.is_cuda and x.shape == (8, 32, 32, 32)
    out = torch.zeros_like(tv)
    g.fill_patches_in_volume(x, out)
    assert torch.equal(out, tv)


def test_empty_and_shape_errors():
    vol = seeded_volume((64, 64, 64), np.uint8, 13)
    p = Patches(vol.shape, "data", points=np.zeros((0, 3), int), widths=np.zeros((0, 3), int))
    assert p.extract(vol, (32,) * 3).shape == (0, 32, 32, 32)
    g = Grid(vol.shape, width=32)
    with pytest.raises(AssertionError):
        g.extract(vol[:32])
    with pytest.raises(AssertionError):
        g.fill_patches_in_volume(np.zeros((7, 32, 32, 32), np.uint8), vol.copy())


def test_scatter_casts_like_numpy_assignment():
    vs = (32, 32, 32)
    g = Grid(vs, width=16)
    sub = np.random.default_rng(5).uniform(0, 3, (len(g), 16, 16, 16)).astype(np.float32)
    out = np.zeros(vs, np.uint8)
    expect = np.zeros(vs, np.uint8)
    g.fill_patches_in_volume(sub, out)
    S.fill_patches_in_volume(sub, g.points, np.full_like(g.points, 16), expect)
    np.testing.assert_array_equal(out, expect)


def test_multiple_grids_scatter_order():
    vs = (70, 66, 80)
    p = Patches(vs, "multiple-grids", min_patch_size=(32,) * 3, max_stride=2, n_points=None)
    sub = [seeded_volume(tuple(int(v) for v in w), np.uint8, 200 + i) for i, w in enumerate(p.widths)]
    out = np.zeros(vs, np.uint8)
    expect = np.zeros(vs, np.uint8)
    p.fill_patches_in_volume(sub, out)
    S.fill_patches_in_volume(sub, p.points, p.widths, expect)
    np.testing.assert_array_equal(out, expect)


def test_config1_scale_gather_properties():
    """BASELINE config 1 sizes: 4096 x 32^3 patches from a 256^3 uint8 volume.  Checked through
    size-independent properties: checksum of checksums against a strided-view oracle, and
    idempotence of fill(extract())."""
    vol = seeded_volume((256,) * 3, np.uint8, 21)
    pts = np.stack([np.random.default_rng(3).integers(0, 256 - 32, 4096) for _ in range(3)], axis=1)
    p = Patches(vol.shape, "data", points=pts, widths=np.full_like(pts, 32))
    out = p.extract(vol, (32,) * 3)
    win = np.lib.stride_tricks.sliding_window_view(vol, (32, 32, 32))
    sums = out.reshape(4096, -1).astype(np.uint64).sum(axis=1)
    expect = np.array([win[z, y, x].astype(np.uint64).sum() for z, y, x in pts])
    np.testing.assert_array_equal(sums, expect)
    idx = np.random.default_rng(0).integers(0, 4096, 64)
    for i in idx:
        z, y, x = pts[i]
        np.testing.assert_array_equal(out[i], vol[z:z + 32, y:y + 32, x:x + 32])
    back = vol.copy()
    p.fill_patches_in_volume(out, back)           # writing a volume's own patches back changes nothing
    np.testing.assert_array_equal(back, vol)


def test_find_min_max_matches_oracle():
    for dt in (np.uint8, np.uint16, np.float16, np.float32):
        vol = seeded_volume((50, 61, 67), dt, 14)
        for s in (1, 2, 4):
            lo, hi = voxel_processing._find_min_max(vol, s)
            elo, ehi = V.find_min_max(vol, s)
            assert lo == float(elo) and hi == float(ehi)
            if dt != np.uint16:
                # pinned host tensor (the streaming path's calling convention): vol[::s, ::s, :] goes up as one strided DMA
                import torch
                pinned = torch.from_numpy(vol).pin_memory()
                assert voxel_processing._find_min_max(pinned, s) == (lo, hi)
                assert voxel_processing._find_min_max(torch.from_numpy(vol).cuda(), s) == (lo, hi)


# ------------------------------------------------------------------ round-2 additions (ADVICE.md findings)
@pytest.mark.parametrize("dt", [np.uint8, np.float32])
def test_2d_binned_extract_matches_numpy_slicing(dt):
    """2-D tables are single-slice volumes: the binning of the image axes must survive the z padding
    (reference: patches.py:748-781 works on any ndim through slices)."""
    img = seeded_volume((1, 96, 128), dt, 300)[0]
    pts = np.array([[0, 0], [3, 5], [32, 64], [17, 33]], dtype=np.uint32)
    wds = np.array([[64, 64], [32, 32], [64, 64], [32, 32]], dtype=np.uint32)
    p = Patches(img.shape, "data", points=pts, widths=wds)
    out = p.extract(img, (32, 32))
    assert out.shape == (4, 32, 32) and out.dtype == img.dtype
    for i in range(4):
        b = int(wds[i, 0]) // 32
        y, x = int(pts[i, 0]), int(pts[i, 1])
        np.testing.assert_array_equal(out[i], img[y:y + wds[i, 0]:b, x:x + wds[i, 1]:b])
    # 2-D scatter of unbinned patches (fill is assignment of full-width patches)
    p1 = Patches(img.shape, "data", points=pts[[1, 3]], widths=wds[[1, 3]])
    sub = p1.extract(img, (32, 32))
    tgt = np.zeros_like(img)
    p1.fill_patches_in_volume(sub, tgt)
    exp = np.zeros_like(img)
    for i in (1, 3):
        y, x = int(pts[i, 0]), int(pts[i, 1])
        exp[y:y + 32, x:x + 32] = img[y:y + 32, x:x + 32]
    np.testing.assert_array_equal(tgt, exp)


def test_variable_width_device_round_trip():
    """extract(vol_cuda, None) returns a list of CUDA tensors (device in -> device out); feeding that list back into
    fill_patches_in_volume must work (patches.py:774-781, 811-821)."""
    vol = seeded_volume(SVS, np.uint16, 301)
    pts = np.array([[0, 0, 0], [8, 8, 16], [20, 4, 0], [1, 20, 30]], dtype=np.uint32)
    wds = np.array([[16, 16, 16], [16, 16, 16], [8, 8, 8], [8, 8, 8]], dtype=np.uint32)
    p = Patches(SVS, "data", points=pts, widths=wds)
    subs = p.extract(torch.from_numpy(vol.astype(np.int32)).to(torch.uint16).cuda(), None)
    assert isinstance(subs, list) and all(t.is_cuda for t in subs)
    out = torch.zeros(SVS, dtype=torch.uint16, device="cuda")
    p.fill_patches_in_volume(subs, out)
    exp = np.zeros_like(vol)
    S.fill_patches_in_volume([vol[z:z + w, y:y + w, x:x + w] for (z, y, x), w in zip(pts.astype(int), wds[:, 0].astype(int))],
                             pts, wds, exp)
    np.testing.assert_array_equal(out.cpu().numpy().view(np.uint16), exp)
    # host lists of arrays still work
    out_h = np.zeros_like(vol)
    p.fill_patches_in_volume([t.cpu().numpy().view(np.uint16) for t in subs], out_h)
    np.testing.assert_array_equal(out_h, exp)


def test_float64_volume_unbinned_gather_scatter_and_refusal():
    """numpy's default dtype: unbinned extract / fill move 64-bit elements as pairs of 32-bit words; a decimating extract
    is refused with one clear message up front."""
    g = np.random.default_rng(302)
    vol = g.normal(size=SVS)
    p = Patches(SVS, "data", points=np.array([[1, 3, 5], [24, 28, 31], [0, 0, 0]], dtype=np.uint32),
                widths=np.full((3, 3), 16, dtype=np.uint32))
    x = p.extract(vol, (16,) * 3)
    assert x.dtype == np.float64
    np.testing.assert_array_equal(x, S.extract(vol, p.points, p.widths, (16,) * 3))
    out = np.zeros_like(vol)
    p.fill_patches_in_volume(x, out)
    exp = np.zeros_like(vol)
    S.fill_patches_in_volume(x, p.points, p.widths, exp)
    np.testing.assert_array_equal(out, exp)
    pb = Patches(SVS, "data", points=np.zeros((1, 3), dtype=np.uint32), widths=np.full((1, 3), 32, dtype=np.uint32))
    with pytest.raises(TypeError, match="binned extract"):
        pb.extract(vol, (16,) * 3)


@pytest.mark.parametrize("dt", [np.uint8, np.float16])
def test_many_overlapping_random_patches_take_the_sorted_processing_order(dt):
    """n P^3 >> volume: the host mirror hands te_gather_perm a (z, y)-sorted processing order (L2 reuse); the output is
    still in patch-index order, bit-exact."""
    vs = (64, 72, 80)
    vol = seeded_volume(vs, dt, 310)
    g = np.random.default_rng(311)
    pts = np.stack([g.integers(0, vs[a] - 16, 3001) for a in range(3)], axis=1).astype(np.uint32)
    p = Patches(vs, "data", points=pts, widths=np.full_like(pts, 16))
    out = p.extract(vol, (16,) * 3)
    np.testing.assert_array_equal(_bits(out), _bits(S.extract(vol, p.points, p.widths, (16,) * 3)))
    tv = torch.from_numpy(vol.view(np.uint8 if dt == np.uint8 else np.int16)).cuda()
    out_d = p.extract(tv, (16,) * 3)
    assert out_d.is_cuda and np.array_equal(out_d.cpu().numpy().view(out.dtype), out)
