"""CPU: the labelling oracle (oracle/voids_np.py) against the golden vectors produced by the reference's own Voids class
(tests/golden/voids_golden.npz, make_voids_golden.py), against a from-scratch union-find restatement of scipy's label
numbering, and against the live reference when /root/reference is mounted."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import ref_loader, voids_np as V


@pytest.fixture(scope="module")
def vg():
    return dict(np.load(os.path.join(GOLDEN, "voids_golden.npz")))


def unpack(prefix, g):
    boxes = g[prefix + "boxes"]
    nv = (boxes[:, 3:] - boxes[:, :3]).prod(axis=1)
    offs = np.concatenate([[0], np.cumsum(nv)])
    masks = [g[prefix + "masks"][offs[i]:offs[i + 1]].reshape(tuple(boxes[i, 3:] - boxes[i, :3])) for i in range(len(boxes))]
    return dict(sizes=g[prefix + "sizes"], cents=g[prefix + "cents"], cpts=g[prefix + "cpts"], boxes=boxes, masks=masks,
                boundary=g[prefix + "boundary"])


def check_voids(v, ref):
    """v: oracle / product dict-like with sizes, cents, cpts, s_voids, x_voids, x_boundary."""
    np.testing.assert_array_equal(np.asarray(v["sizes"], dtype=np.int64), ref["sizes"])
    np.testing.assert_array_equal(np.asarray(v["cents"], dtype=np.int64).reshape(-1, 3), ref["cents"])
    np.testing.assert_array_equal(np.asarray(v["cpts"], dtype=np.int64).reshape(-1, 3), ref["cpts"])
    boxes = np.asarray([[s[0].start, s[1].start, s[2].start, s[0].stop, s[1].stop, s[2].stop] for s in v["s_voids"]]).reshape(-1, 6)
    np.testing.assert_array_equal(boxes, ref["boxes"])
    assert len(v["x_voids"]) == len(ref["masks"])
    for a, b in zip(v["x_voids"], ref["masks"]):
        np.testing.assert_array_equal(np.asarray(a, dtype=np.uint8), b)
    np.testing.assert_array_equal(np.asarray(v["x_boundary"], dtype=np.uint8), ref["boundary"])


CASES = [(2, 0, "cv_d2_"), (0, 0, "cv_d0_"), (1, 2, "cv_d1p2_")]


def test_label_numbering_matches_golden_and_union_find(vg):
    vol = V.void_phantom((48, 40, 56), 3)
    for conn, tag in ((26, "c26_"), (6, "c6_")):
        lab, n = V.label(vol, conn)
        assert n == int(vg[tag + "n"])
        np.testing.assert_array_equal(lab, vg[tag + "labels"])
    small = V.void_phantom((14, 12, 16), 9, p_speck=0.05, n_blobs=6)
    for conn in (6, 26):
        lab, n = V.label(small, conn)
        lab2, n2 = V.label_restated(small, conn)
        assert n == n2
        np.testing.assert_array_equal(lab, lab2)


@pytest.mark.parametrize("dust,pad,tag", CASES)
def test_count_voids_and_export_grid_match_golden(vg, dust, pad, tag):
    v = V.count_voids(vg["c26_labels"], 2, dust, pad_bb=pad)
    check_voids(v, unpack(tag, vg))
    pts, wd, shape, r = V.export_grid(v, 16)
    np.testing.assert_array_equal(pts, vg[tag + "grid_points"])
    assert wd == int(vg[tag + "grid_wd"]) and tuple(shape) == tuple(vg[tag + "grid_shape"]) and r == float(vg[tag + "grid_r"])


def test_import_from_grid_matches_golden(vg):
    from oracle import structures_np as S
    Vf = V.void_phantom((64, 64, 96), 5, p_speck=0.003)
    labb, _ = V.label(Vf[::2, ::2, ::2].copy(), 26)
    vb = V.count_voids(labb, 2, 2)
    check_voids(vb, unpack("ig_b_", vg))
    pts, wd, shape, _ = V.export_grid(vb, 16)
    np.testing.assert_array_equal(pts, vg["ig_grid_points"])
    widths = np.full_like(pts, wd)
    x_grid = S.extract(Vf, pts, widths, (wd,) * 3)
    Vp = np.zeros(shape, dtype=np.uint8)
    S.fill_patches_in_volume(x_grid, pts, widths, Vp)
    vf = V.import_from_grid(vb, Vp)
    vf["x_boundary"] = []
    ref = unpack("ig_f_", vg)
    check_voids(vf, ref)


@pytest.mark.skipif(not ref_loader.available(), reason="needs /root/reference")
def test_oracle_matches_live_reference():
    ref = ref_loader.load()
    vol = V.void_phantom((32, 48, 40), 17, p_speck=0.02, n_blobs=12)
    lab, n = V.label(vol, 26)
    rv = ref.RefVoids().count_voids(lab.copy(), 1, 1)
    ov = V.count_voids(lab, 1, 1)
    np.testing.assert_array_equal(np.asarray(rv["sizes"]), ov["sizes"])
    np.testing.assert_array_equal(np.asarray(rv["cpts"]), ov["cpts"])
    for a, b in zip(rv["x_voids"], ov["x_voids"]):
        np.testing.assert_array_equal(a, b)
    p, r = rv.export_grid(8)
    pts, wd, shape, r2 = V.export_grid(ov, 8)
    np.testing.assert_array_equal(np.asarray(p.points), pts)
    assert r == r2 and p.wd == wd
