"""CPU: the numpy oracle (oracle/structures_np.py, oracle/voxel_np.py) against the golden vectors
generated from the reference's own classes, and against the reference itself when it is mounted."""
import numpy as np
import pytest

from conftest import seeded_volume
from oracle import ref_loader, structures_np as S, voxel_np as V

VS = (351, 350, 340)
SVS = (40, 44, 48)
DTYPES = (("u8", np.uint8), ("u16", np.uint16), ("f16", np.float16), ("f32", np.float32))


def test_known_answers_from_reference_notebooks(golden):
    # SURVEY.md section 4 table (outputs recorded in the reference's notebooks)
    p2, _ = S.multiple_grids_points(VS, (64,) * 3, 2)
    assert p2.shape == (243, 3)
    p4, w4 = S.multiple_grids_points(VS, (64,) * 3, 4)
    assert p4.shape == (259, 3)
    assert p4[:10].tolist() == [[0, 0, 0], [0, 0, 55], [0, 0, 110], [0, 0, 165], [0, 0, 220], [0, 0, 275],
                                [0, 57, 0], [0, 57, 55], [0, 57, 110], [0, 57, 165]]
    g2, _ = S.grid_points(VS, (64,) * 3, 2)
    assert g2.shape == (27, 3)
    gp, _ = S.regular_grid_points((256,) * 3, (32,) * 3)
    assert gp.shape == (512, 3) and gp.dtype == np.uint32


def test_coordinates_match_golden(golden):
    p2, _ = S.multiple_grids_points(VS, (64,) * 3, 2)
    np.testing.assert_array_equal(p2, golden["mg2_points"])
    p4, w4 = S.multiple_grids_points(VS, (64,) * 3, 4)
    np.testing.assert_array_equal(p4, golden["mg4_points"])
    np.testing.assert_array_equal(w4, golden["mg4_widths"])
    g2, gw = S.grid_points(VS, (64,) * 3, 2)
    np.testing.assert_array_equal(g2, golden["grid_s2_points"])
    np.testing.assert_array_equal(gw, golden["grid_s2_widths"])
    rg, _ = S.regular_grid_points(VS, (64,) * 3)
    np.testing.assert_array_equal(rg, golden["rg_points"])
    gp, _ = S.regular_grid_points((256,) * 3, (32,) * 3)
    np.testing.assert_array_equal(gp, golden["grid256_points"])


@pytest.mark.parametrize("name,dt", DTYPES)
def test_gather_scatter_match_golden(golden, name, dt):
    vol = seeded_volume(SVS, dt, 100)
    pts = golden["gs_%s_grid_points" % name]
    wds = np.full_like(pts, 16)
    out = S.extract(vol, pts, wds, (16,) * 3)
    assert out.dtype == vol.dtype
    np.testing.assert_array_equal(out, golden["gs_%s_grid_extract" % name])
    outb = S.extract(vol, golden["gs_%s_bin_points" % name], golden["gs_%s_bin_widths" % name], (16,) * 3)
    np.testing.assert_array_equal(outb, golden["gs_%s_bin_extract" % name])
    sub = seeded_volume((len(pts),) + (16,) * 3, dt, 101)
    vout = seeded_volume(SVS, dt, 102)
    S.fill_patches_in_volume(sub, pts, wds, vout)
    np.testing.assert_array_equal(vout, golden["gs_%s_fill" % name])


def test_binned_extract_is_decimation():
    vol = seeded_volume((320, 300, 310), np.uint8, 7)
    out = S.extract(vol, [[50, 30, 50]], [[256] * 3], (64,) * 3)
    np.testing.assert_array_equal(out[0], vol[50:306:4, 30:286:4, 50:306:4])


def test_regular_grid_round_trip_is_identity():
    vol = seeded_volume((130, 140, 150), np.uint8, 8)
    pts, wds = S.regular_grid_points(vol.shape, (64,) * 3)
    sub = S.extract(vol, pts, wds, (64,) * 3)
    out = np.zeros_like(vol)
    S.fill_patches_in_volume(sub, pts, wds, out)
    np.testing.assert_array_equal(out[:128, :128, :128], vol[:128, :128, :128])
    assert not out[128:].any()


def test_binning_errors():
    with pytest.raises(ValueError):
        S.calc_binning(np.array([[64, 64, 128]]), (32, 32, 32))
    with pytest.raises(ValueError):
        S.calc_binning(np.array([[16, 16, 16]]), (32, 32, 32))
    with pytest.raises(ValueError):
        S.check_valid_points([[0, 0, 300]], [[64] * 3], VS)


def test_voxel_processing_matches_golden(golden):
    vv = seeded_volume((40, 44, 48), np.float32, 104)
    mn, mx = V.find_min_max(vv, 4)
    np.testing.assert_array_equal(np.array([mn, mx], dtype=np.float64), golden["minmax_f32_s4"])
    np.testing.assert_array_equal(V.rescale_data(vv[:4, :4, :4], mn, mx), golden["rescale_f32"])


@pytest.mark.skipif(not ref_loader.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_against_live_reference():
    ref = ref_loader.load()
    vol = seeded_volume((70, 90, 110), np.uint16, 9)
    for mode, kw in (("grid", dict(patch_size=(32,) * 3, stride=1)), ("grid", dict(patch_size=(16,) * 3, stride=2)),
                     ("regular-grid", dict(patch_size=(32,) * 3))):
        rp = ref.RefPatches(vol.shape, mode, **kw)
        if mode == "grid":
            pts, wds = S.grid_points(vol.shape, kw["patch_size"], kw["stride"])
        else:
            pts, wds = S.regular_grid_points(vol.shape, kw["patch_size"])
        np.testing.assert_array_equal(pts, rp.points)
        np.testing.assert_array_equal(wds, rp.widths)
        ps = kw["patch_size"]
        np.testing.assert_array_equal(S.extract(vol, pts, wds, ps), rp.extract(vol, ps))
