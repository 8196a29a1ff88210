"""CPU: host-side coordinate logic of tomoencoders_b200.Patches / Grid against the golden vectors
(generated from the reference classes) and the reference's error behaviour.  No GPU calls."""
import numpy as np
import pytest

from tomoencoders_b200 import Grid, Patches

VS = (351, 350, 340)


def test_initialisers_match_reference(golden):
    p2 = Patches(VS, "multiple-grids", min_patch_size=(64,) * 3, max_stride=2, n_points=None)
    assert p2.points.dtype == np.uint32 and p2.widths.dtype == np.uint32
    np.testing.assert_array_equal(p2.points, golden["mg2_points"])
    p4 = Patches(VS, "multiple-grids", min_patch_size=(64,) * 3, max_stride=4, n_points=None)
    np.testing.assert_array_equal(p4.points, golden["mg4_points"])
    np.testing.assert_array_equal(p4.widths, golden["mg4_widths"])
    g2 = Patches(VS, "grid", patch_size=(64,) * 3, stride=2)
    np.testing.assert_array_equal(g2.points, golden["grid_s2_points"])
    np.testing.assert_array_equal(g2.widths, golden["grid_s2_widths"])
    assert len(g2) == 27
    rg = Patches(VS, "regular-grid", patch_size=(64,) * 3)
    np.testing.assert_array_equal(rg.points, golden["rg_points"])
    oc = Patches(VS, "octree-grid", patch_size=(64,) * 3)
    np.testing.assert_array_equal(oc.points, rg.points)
    g = Grid((256,) * 3, width=32)
    np.testing.assert_array_equal(g.points, golden["grid256_points"])
    assert g.points.dtype == np.uint32 and g.wd == 32 and g["width"] == 32 and len(g) == 512


def test_dict_interface_and_slices():
    p = Patches(VS, "data", points=[[0, 10, 50], [50, 30, 50]], widths=[[64] * 3, [256] * 3])
    assert p["source"] == "data" and p["points"] is p.points and p["widths"] is p.widths
    s = p.slices()
    assert s.tolist() == [[slice(0, 64, 1), slice(10, 74, 1), slice(50, 114, 1)],
                          [slice(50, 306, 1), slice(30, 286, 1), slice(50, 306, 1)]]
    assert p._calc_binning((64,) * 3).tolist() == [1, 4]
    assert p.centers().tolist() == [[32, 42, 82], [178, 158, 178]]


def test_random_initialisers_follow_legacy_rng():
    np.random.seed(5)
    p = Patches((256,) * 3, "random-fixed-width", patch_size=(32,) * 3, n_points=100)
    np.random.seed(5)
    expect = np.asarray([np.random.randint(0, 256 - 32, 100) for _ in range(3)]).T
    np.testing.assert_array_equal(p.points, expect)
    assert p.points.max() < 224            # the last valid corner is never drawn (high exclusive)
    np.random.seed(6)
    q = Patches((256,) * 3, "random", min_patch_size=(32,) * 3, max_stride=4, n_points=50)
    assert set(np.unique(q.widths)) <= {32, 64, 96}
    assert np.all(q.points.astype(np.int64) + q.widths <= 256)
    with pytest.raises(ValueError):
        Patches((256,) * 3, "random", min_patch_size=(32,) * 3, max_stride=1, n_points=5)


def test_invalid_points_and_arguments_raise():
    with pytest.raises(ValueError):
        Patches(VS, "data", points=[[0, 0, 300]], widths=[[64] * 3])
    with pytest.raises(ValueError):
        Patches(VS, "data", points=[[-5, 0, 0]], widths=[[64] * 3])       # would wrap through uint32
    with pytest.raises(ValueError):
        Patches(VS, "data", points=[[0, 0, 0], [1, 1, 1]], widths=[[64] * 3])
    with pytest.raises(ValueError):
        Patches(VS, "grid", patch_size=(64,) * 3, stride=6)
    with pytest.raises(NotImplementedError):
        Patches(VS, "no-such-mode")
    with pytest.raises(AssertionError):
        Grid((250,) * 3, width=32)
    p = Patches(VS, "data", points=[[0, 0, 0]], widths=[[64, 64, 128]])
    with pytest.raises(ValueError):
        p._calc_binning((32,) * 3)
    with pytest.raises(ValueError):
        Patches(VS, "data", points=[[0, 0, 0]], widths=[[16] * 3])._calc_binning((32,) * 3)


def test_features_filter_sort_select():
    p = Patches(VS, "regular-grid", patch_size=(64,) * 3)
    n = len(p)
    f = np.arange(n, dtype=np.float32)[::-1].copy()
    p.add_features(f, names=["score"])
    p.add_features(np.stack([f, 2 * f], axis=1))
    assert p.feature_names == ["score", "Unnamed_f0", "Unnamed_f1"] and p.features.shape == (n, 3)
    np.testing.assert_array_equal(p["score"], f)
    with pytest.raises(ValueError):
        p.add_features(np.zeros(n + 1))
    top = p.select_by_feature(5, ife=0, selection_by="highest")
    np.testing.assert_array_equal(top.features[:, 0], f[:5])
    low = p.select_by_feature(5, ife=0, selection_by="lowest")
    np.testing.assert_array_equal(low.features[:, 0], f[::-1][:5])
    srt = p.sort_by_feature(ife=0)
    assert np.all(np.diff(srt.features[:, 0]) >= 0)
    cond = np.stack([f > 10, f < 50], axis=1)
    flt = p.filter_by_condition(cond)
    assert len(flt) == int(np.sum((f > 10) & (f < 50)))
    with pytest.raises(ValueError):
        p.filter_by_condition(np.ones(n + 1))
    assert len(p.select_by_range((3, 9))) == 6 and len(p.pop(2)) == n - 2 and len(p.pop(-3)) == n - 3
    pl = p.select_by_plane(0, 70)
    assert np.all(pl.points[:, 0] == 64)
    r = p.rescale(2, tuple(2 * v for v in VS))
    np.testing.assert_array_equal(r.points, p.points * 2)
    with pytest.raises(ValueError):
        p.rescale(2, (100, 100, 100))
    c = p.copy()
    c.append(p)
    assert len(c) == 2 * n and c.features.shape == (2 * n, 3)


def test_perturb_keeps_points_valid():
    # /root/reference/tests/patches/test_perturb.py:1-16 (regular grid, cylindrical filter, perturb)
    p = Patches((1000, 2100, 2100), "regular-grid", patch_size=(64,) * 3)
    p = p.filter_by_cylindrical_mask(mask_ratio=0.98)
    np.random.seed(0)
    q = p.perturb(5)
    q._check_valid_points()
    assert len(q) == len(p) and np.abs(q.points.astype(np.int64) - p.points.astype(np.int64)).max() <= 5


def test_million_patches_is_fast():
    import time
    t0 = time.time()
    np.random.seed(1)
    p = Patches((2048,) * 3, "random-fixed-width", patch_size=(64,) * 3, n_points=1_000_000)
    assert len(p) == 1_000_000 and time.time() - t0 < 5.0


def test_grid_methods_and_dump_load(tmp_path):
    g = Grid((128, 256, 256), width=64)
    assert len(g) == 2 * 4 * 4
    assert len(g.filter_by_cylindrical_mask(0.5)) < len(g)
    c = g.copy()
    np.testing.assert_array_equal(c.points, g.points)
    r = g.rescale(2)
    assert r.vol_shape == (256, 512, 512) and r.wd == 128
    assert g.slices()[1].tolist() == [slice(0, 64), slice(0, 64), slice(64, 128)]
    f = str(tmp_path / "g.npz")
    g.dump(f)
    g2 = Grid(None if False else (1, 1, 1), "file", fpath=f)
    np.testing.assert_array_equal(g2.points, g.points)
    assert g2.vol_shape == g.vol_shape and g2.wd == 64
    p = Patches((128, 256, 256), "regular-grid", patch_size=(64,) * 3)
    p.add_features(np.arange(len(p), dtype=np.float32), names=["a"])
    f = str(tmp_path / "p.npz")
    p.dump(f)
    p2 = Patches((1, 1, 1), "file", fpath=f)
    np.testing.assert_array_equal(p2.points, p.points)
    assert p2.feature_names == ["a"] and p2.vol_shape == p.vol_shape
