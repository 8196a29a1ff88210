"""GPU: labelling / Voids through the reference-facing API (tomoencoders_b200.structures.voids -> C ABI -> CUDA) against
the golden vectors of the reference's own Voids class and against the oracle on seeded volumes.  Integer work: bit-exact."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from oracle import voids_np as V
from test_oracle_voids import CASES, check_voids, unpack
from tomoencoders_b200 import Grid
from tomoencoders_b200.structures.voids import Voids, label

pytestmark = pytest.mark.gpu

ONES = np.ones((3, 3, 3), dtype=np.uint8)


@pytest.fixture(scope="module")
def vg():
    return dict(np.load(os.path.join(GOLDEN, "voids_golden.npz")))


def test_label_matches_golden(vg):
    vol = V.void_phantom((48, 40, 56), 3)
    for structure, tag in ((ONES, "c26_"), (None, "c6_")):
        lab, n = label(vol, structure=structure)
        assert isinstance(lab, np.ndarray) and lab.dtype == np.int32 and n == int(vg[tag + "n"])
        np.testing.assert_array_equal(lab, vg[tag + "labels"])
    # device in -> device out, other input dtypes
    lab_d, n = label(torch.from_numpy(vol.astype(np.float32)).cuda(), structure=ONES)
    assert lab_d.is_cuda and n == int(vg["c26_n"])
    np.testing.assert_array_equal(lab_d.cpu().numpy(), vg["c26_labels"])


@pytest.mark.parametrize("shape,seed,p", [((33, 47, 65), 1, 0.02), ((64, 64, 64), 2, 0.3), ((5, 7, 300), 3, 0.5), ((1, 1, 9), 4, 0.5),
                                          ((96, 128, 160), 5, 0.0)])
def test_label_matches_oracle_on_seeded_volumes(shape, seed, p):
    """Dense random noise (p = 0.3 / 0.5: percolating, worst case for union-find), ragged shapes, degenerate axes."""
    g = np.random.default_rng(seed)
    vol = (g.random(shape) < p).astype(np.uint8) if p > 0 else V.void_phantom(shape, seed, n_blobs=60)
    for conn, structure in ((26, ONES), (6, None)):
        ref, n_ref = V.label(vol, conn)
        lab, n = label(vol, structure=structure)
        assert n == n_ref
        np.testing.assert_array_equal(lab, ref)


def test_label_edge_cases():
    lab, n = label(np.zeros((8, 8, 8), np.uint8), structure=ONES)
    assert n == 0 and not lab.any()
    lab, n = label(np.ones((8, 9, 10), np.uint8))
    assert n == 1 and (lab == 1).all()
    with pytest.raises(NotImplementedError):
        label(np.ones((3, 3, 3), np.uint8), structure=np.eye(3)[None].repeat(3, 0))


@pytest.mark.parametrize("dust,pad,tag", CASES)
def test_count_voids_and_export_grid_match_golden(vg, dust, pad, tag):
    v = Voids()
    v.pad_bb = pad
    v.count_voids(vg["c26_labels"], 2, dust)
    check_voids(v, unpack(tag, vg))
    assert v.b == 2 and v.n_voids == len(v["sizes"])
    p, r = v.export_grid(16)
    assert isinstance(p, Grid)
    np.testing.assert_array_equal(np.asarray(p.points), vg[tag + "grid_points"])
    assert p.wd == int(vg[tag + "grid_wd"]) and tuple(p.vol_shape) == tuple(vg[tag + "grid_shape"]) and r == float(vg[tag + "grid_r"])
    # the painted volume itself (overlapping boxes: later voids overwrite earlier ones) against the oracle
    np.testing.assert_array_equal(v.paint().cpu().numpy(), V.paint(V.count_voids(vg["c26_labels"], 2, dust, pad_bb=pad)))


def test_import_from_grid_matches_golden(vg):
    Vf = V.void_phantom((64, 64, 96), 5, p_speck=0.003)
    labb, _ = label(Vf[::2, ::2, ::2].copy(), structure=ONES)
    vb = Voids().count_voids(labb, 2, 2)
    check_voids(vb, unpack("ig_b_", vg))
    p_grid, _ = vb.export_grid(16)
    np.testing.assert_array_equal(np.asarray(p_grid.points), vg["ig_grid_points"])
    x_grid = p_grid.extract(Vf)
    vf = Voids().import_from_grid(vb, x_grid, p_grid)
    check_voids(vf, unpack("ig_f_", vg))
    assert vf.b == 1 and tuple(vf.vol_shape) == (64, 64, 96)


def test_selection_helpers_follow_the_reference(vg):
    v = Voids().count_voids(vg["c26_labels"], 2, 0)
    n0 = len(v)
    v.sort_by_size(reverse=True)
    assert np.all(np.diff(v["sizes"]) <= 0) and len(v) == n0
    v.select_by_size(4.0, pixel_size_um=1.0)                      # threshold (4 / b)^3 = 8 voxels
    assert len(v) < n0 and np.all(v["sizes"] >= 8)
    v.select_around_void(0, 40.0)
    assert "distance_0" in v and len(v) >= 1
    c0 = np.asarray(v["cpts"]).copy()
    v.transform_linear_shift((1, 2, 3))
    np.testing.assert_array_equal(np.asarray(v["cpts"]), c0 + np.asarray((1, 2, 3)))


def test_coarse_to_fine_round_trip_at_512_cubed():
    """Size-independent property at a BASELINE-sized mask (512^3): every voxel of a component carries the label of the
    component's first voxel in C order, labels are 1..n without gaps, and relabelling the label image reproduces it."""
    g = torch.Generator(device="cuda").manual_seed(0)
    vol = (torch.rand((512, 512, 512), device="cuda", generator=g) < 0.2).to(torch.uint8)
    lab, n = label(vol, structure=ONES)
    assert lab.is_cuda and n > 0 and int(lab.max()) == n
    assert torch.equal(lab > 0, vol > 0)
    firsts = torch.zeros(n + 1, dtype=torch.int64, device="cuda").fill_(2 ** 62)
    idx = torch.arange(lab.numel(), device="cuda")
    firsts.scatter_reduce_(0, lab.reshape(-1).long(), idx, reduce="amin")
    assert bool((firsts[1:][1:] > firsts[1:][:-1]).all())       # numbered in C order of the first voxel
    lab2, n2 = label(lab, structure=ONES)                       # non-zero = foreground: same components
    assert n2 == n and torch.equal(lab2, lab)
