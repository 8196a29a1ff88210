"""GPU: misc/voxel_processing (te_minmax_nd / te_histogram_nd / te_rescale / te_edge_map /
te_cylindrical_mask / te_cyl_values) bit-exact against the numpy oracle and the golden vectors of the
reference's own functions."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import voxel_np as V
from oracle.phantom import cosine_labels, cosine_volume

pytestmark = pytest.mark.gpu

DTYPES = (("u8", np.uint8), ("u16", np.uint16), ("f16", np.float16), ("f32", np.float32))
SHAPE = (24, 40, 40)


@pytest.fixture(scope="module")
def vp():
    import torch
    assert torch.cuda.is_available()
    from tomoencoders_b200.misc import voxel_processing
    return voxel_processing


@pytest.fixture(scope="module")
def vg():
    return dict(np.load(os.path.join(GOLDEN, "voxel_golden.npz")))


def _dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("name,dt", DTYPES)
def test_autocontrast_matches_reference_golden(vp, vg, name, dt):
    vol = cosine_volume(SHAPE, dt, 200)
    for src in (vol, _dev(vol)):                                 # host in / device in
        for f in (1, 2, 3):
            got = vp.modified_autocontrast(src, s=0.01, normalize_sampling_factor=f)
            np.testing.assert_array_equal(np.array(got, dtype=np.float64), vg["ac_%s_f%d" % (name, f)])
    got = vp.modified_autocontrast(vol, s=(0.05, 0.1), normalize_sampling_factor=2)
    np.testing.assert_array_equal(np.array(got, dtype=np.float64), vg["ac_%s_tuple" % name])
    got = vp.modified_autocontrast(vol.reshape(-1)[:5001], s=0.02)
    np.testing.assert_array_equal(np.array(got, dtype=np.float64), vg["ac_%s_1d" % name])
    v4 = np.stack([cosine_volume((16, 16, 16), dt, 201 + i) for i in range(3)])
    got = vp.modified_autocontrast(v4, s=0.05, normalize_sampling_factor=2)
    np.testing.assert_array_equal(np.array(got, dtype=np.float64), vg["ac_%s_4d" % name])


@pytest.mark.parametrize("name,dt", DTYPES)
@pytest.mark.parametrize("shape,f", [((24, 40, 40), 2), ((17, 33, 67), 1), ((9, 20, 4099), 1), ((5, 7, 11), 3)])
def test_histogram_bit_exact_against_numpy(vp, vg, name, dt, shape, f):
    """Counts of every bin: vectorised path (aligned rows), scalar path (odd extents, strided views),
    rows longer than one block segment."""
    vol = cosine_volume(shape, dt, 300 + f)
    view = _dev(vol)[::f, ::f, ::f]
    h, edges = vp._histogram_view(view, 500)
    hn, en = np.histogram(vol[::f, ::f, ::f].reshape(-1), bins=500)
    np.testing.assert_array_equal(edges, en)
    np.testing.assert_array_equal(h, hn)
    assert h.sum() == view.numel()
    if shape == SHAPE and f == 2:       # the reference-side golden counts (np.histogram run at generation time)
        hg, _ = vp._histogram_view(_dev(cosine_volume(SHAPE, dt, 200))[::2, ::2, ::2], 500)
        np.testing.assert_array_equal(hg, vg["hist_%s" % name])


def test_histogram_of_constant_volume_and_errors(vp):
    import torch
    c = torch.full((8, 16, 16), 7, dtype=torch.uint8, device="cuda")
    h, edges = vp._histogram_view(c, 500)
    hn, en = np.histogram(c.cpu().numpy().reshape(-1), bins=500)
    np.testing.assert_array_equal(h, hn)
    np.testing.assert_array_equal(edges, en)
    with pytest.raises(ValueError):
        vp.modified_autocontrast(torch.zeros((2, 2), device="cuda"))


@pytest.mark.parametrize("name,dt", DTYPES)
@pytest.mark.parametrize("shape", [SHAPE, (7, 18, 18), (5, 32, 32)])
def test_edge_map_and_cylinder_ops_bit_exact(vp, vg, name, dt, shape):
    vol = cosine_volume(shape, dt, 200)
    lab = (vol > vol.mean()).astype(dt)
    e_host = vp.edge_map(lab)
    assert isinstance(e_host, np.ndarray) and e_host.dtype == np.bool_
    np.testing.assert_array_equal(e_host, V.edge_map(lab))
    e_dev = vp.edge_map(_dev(lab))
    assert e_dev.is_cuda
    np.testing.assert_array_equal(e_dev.cpu().numpy(), V.edge_map(lab))
    for fac, val in ((0.8, 3), (1.0, 0), (0.3, 1), (1.5, 2), (0.0, 5)):
        a, b = vol.copy(), vol.copy()
        assert vp.cylindrical_mask(a, fac, mask_val=val) is None
        V.cylindrical_mask(b, fac, mask_val=val)
        np.testing.assert_array_equal(a, b)
        d = _dev(vol)
        vp.cylindrical_mask(d, fac, mask_val=val)
        np.testing.assert_array_equal(d.cpu().numpy(), b)
        np.testing.assert_array_equal(vp.get_values_cyl_mask(vol, fac), V.get_values_cyl_mask(vol, fac))
    if shape == SHAPE:
        np.testing.assert_array_equal(e_host, vg["edge_%s" % name])
        m = vol.copy()
        vp.cylindrical_mask(m, 0.8, mask_val=3)
        np.testing.assert_array_equal(m, vg["cyl_%s" % name])
        np.testing.assert_array_equal(vp.get_values_cyl_mask(vol, 0.9), vg["cylvals_%s" % name])


def test_edge_map_bool_labels_and_unaligned_rows(vp, vg):
    lab = cosine_labels(SHAPE, 210)
    np.testing.assert_array_equal(vp.edge_map(lab.astype(np.uint8)), vg["edge_bool"])
    np.testing.assert_array_equal(vp.edge_map(lab), vg["edge_bool"])                    # numpy bool in
    odd = cosine_labels((9, 13, 21), 211).astype(np.uint8)                              # scalar kernel
    np.testing.assert_array_equal(vp.edge_map(odd), V.edge_map(odd))
    flat = np.zeros((4, 16, 16), dtype=np.uint8)
    assert not vp.edge_map(flat).any()
    i32 = cosine_labels((6, 8, 12), 212).astype(np.int32) * 70000
    np.testing.assert_array_equal(vp.edge_map(i32), V.edge_map(i32))


def test_cylinder_golden_subsampled_view_and_odd_extent(vp, vg):
    m1 = cosine_volume(SHAPE, np.float32, 200)
    vp.cylindrical_mask(m1, 1.0, mask_val=np.min(m1[::2, ::2, ::2]))
    np.testing.assert_array_equal(m1, vg["cyl_f32_full"])
    dv = _dev(cosine_volume(SHAPE, np.float32, 200))
    got = vp.get_values_cyl_mask(dv[::2, ::2, ::2], 1.0)           # strided device view, read in place
    np.testing.assert_array_equal(got.cpu().numpy(), vg["cylvals_f32_sub"])
    with pytest.raises(IndexError):
        vp.cylindrical_mask(np.zeros((4, 9, 9), dtype=np.uint8), 0.5)
    with pytest.raises(AssertionError):
        vp.get_values_cyl_mask(np.zeros((4, 8, 10), dtype=np.uint8), 0.5)


def test_normalize_volume_bit_exact_float32(vp, vg):
    for f in (1, 4):
        nv = cosine_volume(SHAPE, np.float32, 220)
        ret = vp.normalize_volume_gpu(nv, chunk_size=8, normalize_sampling_factor=f)
        assert ret is nv
        np.testing.assert_array_equal(nv, vg["norm_f32_f%d" % f])
        d = _dev(cosine_volume(SHAPE, np.float32, 220))
        assert vp.normalize_volume_gpu(d, normalize_sampling_factor=f) is d
        np.testing.assert_array_equal(d.cpu().numpy(), vg["norm_f32_f%d" % f])
    # chunked two-pass route of host volumes larger than the whole-volume limit
    old = vp._WHOLE_VOLUME_BYTES
    vp._WHOLE_VOLUME_BYTES = 0
    try:
        nv = cosine_volume(SHAPE, np.float32, 220)
        vp.normalize_volume_gpu(nv, chunk_size=7, normalize_sampling_factor=4)
        np.testing.assert_array_equal(nv, vg["norm_f32_f4"])
    finally:
        vp._WHOLE_VOLUME_BYTES = old
    with pytest.raises(TypeError):
        vp.normalize_volume_gpu(np.zeros((4, 4, 4), dtype=np.uint8))
    h = cosine_volume(SHAPE, np.float16, 221)
    want = V.rescale_data(h.astype(np.float32), np.float32(h.min()), np.float32(h.max())).astype(np.float16)
    vp.normalize_volume_gpu(h)
    np.testing.assert_array_equal(h, want)


@pytest.mark.parametrize("name,dt", DTYPES)
def test_rescale_data_device(vp, name, dt):
    x = cosine_volume((6, 10, 19), dt, 230)
    lo, hi = float(x.min()), float(x.max())
    got = vp._rescale_data(_dev(x), lo, hi).cpu().numpy()
    want = ((x.astype(np.float32) - np.float32(lo)) / np.float32(np.float32(hi) - np.float32(lo) + np.float32(1e-12)))
    if dt == np.float16:
        want = want.astype(np.float16)
    np.testing.assert_array_equal(got, want.astype(got.dtype))
    # host arrays keep the reference's numpy arithmetic
    np.testing.assert_array_equal(vp._rescale_data(x, lo, hi), V.rescale_data(x, lo, hi))


def test_full_size_properties(vp):
    """BASELINE-size volume (512^3): size-independent properties."""
    import torch
    g = torch.Generator(device="cuda").manual_seed(3)
    vol = torch.rand((512, 512, 512), generator=g, device="cuda").to(torch.float16)
    h, edges = vp._histogram_view(vol[::2, ::2, ::2], 500)
    assert int(h.sum()) == 256 ** 3
    lo, hi = vp.modified_autocontrast(vol, s=0.01, normalize_sampling_factor=2)
    assert 0.0 <= lo < 0.05 and 0.95 < hi <= 1.0
    lab = (vol > 0.5).to(torch.uint8)
    e = vp.edge_map(lab)
    ref = torch.zeros_like(e)
    for ax in range(3):
        d = lab.narrow(ax, 0, 511) != lab.narrow(ax, 1, 511)
        ref.narrow(ax, 0, 511).logical_or_(d)
        ref.narrow(ax, 1, 511).logical_or_(d)
    assert torch.equal(e, ref)
    m = lab.clone()
    vp.cylindrical_mask(m, 1.0, mask_val=7)
    inside = vp.get_values_cyl_mask(m, 1.0)
    assert not bool((inside == 7).any())
    assert int((m == 7).sum()) == m.numel() - inside.numel()
    assert torch.equal(inside, vp.get_values_cyl_mask(lab, 1.0))

