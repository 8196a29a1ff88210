"""GPU: the training data generator kernels (te_patch_std, te_training_pairs) against the numpy restatement of
keras_processor.py:213-232 / surface_segmenter.py:75-82 -- bit-exact for the deterministic part (binned gather + rot90
for every k and every ordered axis pair), statistical for the counter-based noise."""
import numpy as np
import pytest
import torch

from oracle import train_data_np as TD
from tomoencoders_b200 import Patches
from tomoencoders_b200.neural_nets import SurfaceSegmenter
from tomoencoders_b200.neural_nets import surface_segmenter as SS

pytestmark = pytest.mark.gpu
VS = (72, 80, 96)


def _vols(xdt):
    g = np.random.default_rng(0)
    X = g.normal(0.5, 0.2, VS).astype(xdt) if np.issubdtype(xdt, np.floating) else g.integers(0, 60000, VS).astype(xdt)
    Y = (g.uniform(size=VS) < 0.3).astype(np.uint8)
    Y[:24, :24, :24] = 0                                            # a blank region
    return X, Y


def _patches():
    pts = np.array([[0, 0, 0], [5, 7, 9], [8, 16, 32], [40, 48, 64], [0, 0, 0], [3, 1, 2], [24, 30, 11], [1, 2, 3]], np.uint32)
    wds = np.array([[16] * 3, [16] * 3, [32] * 3, [32] * 3, [48] * 3, [16] * 3, [48] * 3, [64] * 3], np.uint32)
    return Patches(VS, "data", points=pts, widths=wds)


@pytest.mark.parametrize("xdt", [np.float32, np.float16, np.uint16])
def test_training_pairs_deterministic_part_is_bit_exact(xdt):
    X, Y = _vols(xdt)
    p = _patches()
    n = len(p)
    # every (k, ordered axis pair) combination at least once
    combos = [(k, a, b) for k in range(4) for a in range(3) for b in range(3) if a != b]
    for start in range(0, len(combos), n):
        sel = [combos[(start + i) % len(combos)] for i in range(n)]
        nrots = np.array([c[0] for c in sel], np.int32)
        axes = np.array([[c[1], c[2]] for c in sel], np.int32)
        x, y = SS._training_pairs(X, Y, p, (16,) * 3, None, nrots, axes, 1)
        xr, yr = TD.training_pairs(X, Y, p.points, p.widths, (16,) * 3, None, nrots, axes)
        assert x.dtype == torch.float32 and y.dtype == torch.uint8 and tuple(x.shape) == (n, 16, 16, 16, 1)
        np.testing.assert_array_equal(x.cpu().numpy(), xr.astype(np.float32))
        np.testing.assert_array_equal(y.cpu().numpy(), yr)
    # no rotation, zero sigma: the plain binned extract
    x, y = SS._training_pairs(X, Y, p, (16,) * 3, np.zeros(n, np.float32), None, None, 1)
    np.testing.assert_array_equal(x.cpu().numpy()[..., 0], p.extract(X, (16,) * 3).astype(np.float32))


def test_noise_is_reproducible_gaussian_with_the_requested_sigma():
    X, Y = _vols(np.float32)
    p = _patches()
    n = len(p)
    sigma = np.linspace(0.0, 0.35, n).astype(np.float32)
    x0, _ = SS._training_pairs(X, Y, p, (16,) * 3, None, None, None, 7)
    x1, _ = SS._training_pairs(X, Y, p, (16,) * 3, sigma, None, None, 7)
    x2, _ = SS._training_pairs(X, Y, p, (16,) * 3, sigma, None, None, 7)
    x3, _ = SS._training_pairs(X, Y, p, (16,) * 3, sigma, None, None, 8)
    assert torch.equal(x1, x2) and not torch.equal(x1, x3)
    d = (x1 - x0).cpu().numpy().reshape(n, -1).astype(np.float64)
    assert np.all(d[0] == 0.0)
    for i in range(1, n):
        z = d[i] / sigma[i]
        assert abs(z.mean()) < 0.06 and abs(z.std() - 1.0) < 0.05, (i, z.mean(), z.std())
        assert abs((np.abs(z) < 1.0).mean() - 0.6827) < 0.03 and np.abs(z).max() < 6.0
    # patches are decorrelated from each other (the patch index is part of the counter)
    assert abs(np.corrcoef(d[3] / sigma[3], d[4] / sigma[4])[0, 1]) < 0.06


def test_patch_std_and_blank_rejection_match_numpy():
    X, Y = _vols(np.float32)
    p = _patches()
    got = SS._patch_std(Y, p, (16,) * 3).cpu().numpy()
    want = TD.find_blanks_std(Y, p.points, p.widths, (16,) * 3)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-7)
    got_x = SS._patch_std(X, p, (16,) * 3).cpu().numpy()
    np.testing.assert_allclose(got_x, TD.find_blanks_std(X.astype(np.float64), p.points, p.widths, (16,) * 3), rtol=1e-5)
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="t", n_filters=[16], n_blocks=1, activation="lrelu",
                          batch_norm=True, isconcat=[True], pool_size=[2])
    cond = fe._find_blanks(p, 0.2, Y, (16,) * 3)
    np.testing.assert_array_equal(cond, want > want.max() * 0.2)
    assert not cond[0] and cond[1:].any()                           # the patch inside the blank region is rejected


def test_generator_yields_reference_shaped_batches_with_the_reference_draw_order():
    X, Y = _vols(np.float32)
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="t", n_filters=[16], n_blocks=1, activation="lrelu",
                          batch_norm=True, isconcat=[True], pool_size=[2])
    np.random.seed(11)
    p = fe.get_training_patches(VS, 6, (16,) * 3, 3, 0.1, Y)
    np.random.seed(12)
    x, y = fe.extract_training_patch_pairs(X, Y, p, 0.2, True, (16,) * 3)
    # replay the reference's host draws (keras_processor.py:221, 225-227) and compare the deterministic part
    np.random.seed(12)
    std_batch = np.random.uniform(0, 0.2, 6)
    nrots = np.random.randint(0, 4, 6)
    axes = np.stack([np.random.choice([0, 1, 2], size=2, replace=False) for _ in range(6)])
    xr, yr = TD.training_pairs(X, Y, p.points, p.widths, (16,) * 3, None, nrots, axes)
    np.testing.assert_array_equal(y.cpu().numpy(), yr)
    resid = (x.cpu().numpy() - xr.astype(np.float32)).reshape(6, -1)
    np.testing.assert_allclose(resid.std(axis=1), std_batch, rtol=0.08, atol=1e-3)
    xb, yb = next(fe._data_generator([X], [Y], 6, (16,) * 3, 3, 0.1, True, 0.1))
    assert tuple(xb.shape) == (6, 16, 16, 16, 1) == tuple(yb.shape) and xb.dtype == torch.float32 and yb.dtype == torch.uint8
