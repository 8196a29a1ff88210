"""CPU: weight specifications of the product (tomoencoders_b200/neural_nets/_spec.py) against the
oracle's restatement of the Keras builders, model-tag FLOP counts against SURVEY.md section 8d, and
the PCA oracle against scikit-learn."""
import numpy as np
import pytest

from oracle import nets_torch as O, pca_np, weights as OW
from tomoencoders_b200.neural_nets import _spec

UNETS = [
    dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], isconcat=[True] * 3),       # M_a01
    dict(n_filters=[16, 32], n_blocks=2, pool_size=[2, 4], isconcat=[True] * 2),              # M_a02
    dict(n_filters=[8, 16], n_blocks=2, pool_size=[2, 4], isconcat=[True] * 2),               # M_a07
    dict(n_filters=[16, 32], n_blocks=2, pool_size=[2, 2], isconcat=[False, True]),
    dict(n_filters=[16, 32], n_blocks=2, pool_size=[1, 2], isconcat=[True, True], batch_norm=False),
]
ENCS = [
    ((64, 64, 64), dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], hidden_units=[128, 128], POOL_FLAG=True)),
    ((32, 32, 32), dict(n_filters=[16, 32, 64], n_blocks=3, pool_size=[2, 2, 2], hidden_units=[128, 128], POOL_FLAG=True)),
    ((32, 32, 32), dict(n_filters=[32, 64], n_blocks=2, pool_size=2, hidden_units=[128, 32, 2], POOL_FLAG=False, batch_norm=False)),
]


@pytest.mark.parametrize("params", UNETS)
def test_unet_spec_matches_oracle(params):
    assert _spec.unet_spec(**params) == O.unet_spec(**params)


@pytest.mark.parametrize("size,params", ENCS)
def test_encoder_decoder_spec_matches_oracle(size, params):
    spec, flat, pre = O.encoder_spec(size, **params)
    assert _spec.encoder_spec(size, **params) == spec
    assert _spec.encoder_shapes(size, params["n_filters"], params["n_blocks"], params["pool_size"], params["POOL_FLAG"]) == (flat, tuple(pre))
    assert _spec.decoder_spec(size, **params) == O.decoder_spec(size, **params)


def test_keras_default_init_shapes_and_statistics():
    spec = _spec.unet_spec(**UNETS[0])
    w = _spec.keras_default_init(spec, seed=0)
    assert [tuple(a.shape) for a in w] == [s for _, s in spec]
    k = w[0]                                     # (3,3,3,1,16) glorot uniform
    lim = np.sqrt(6.0 / (27 * 1 + 27 * 16))
    assert np.abs(k).max() <= lim and np.abs(k).max() > 0.8 * lim
    kinds = [k for k, _ in spec]
    assert np.all(w[kinds.index("bn_gamma")] == 1) and np.all(w[kinds.index("conv_b")] == 0)
    assert sum(a.size for a in w) == 4_778_977 or sum(a.size for a in w) > 4_000_000


def test_oracle_unet_tiny_forward_is_consistent_in_fp64():
    params = dict(n_filters=[4, 8], n_blocks=2, pool_size=[2, 2], isconcat=[True, True])
    w = OW.draw(O.unet_spec(**params), 3)
    x = np.random.default_rng(0).uniform(0, 1, (2, 16, 16, 16, 1)).astype(np.float32)
    import torch
    y32 = O.unet_forward(x, w, **params)
    y64 = O.unet_forward(x, w, dtype=torch.float64, **params)
    assert y32.shape == x.shape and np.abs(y32 - y64).max() < 1e-5
    # transposed conv with k=2 < stride=4 leaves bias-only gaps
    params4 = dict(n_filters=[4], n_blocks=1, pool_size=[4], isconcat=[True])
    w4 = OW.draw(O.unet_spec(**params4), 4)
    assert O.unet_forward(x, w4, **params4).shape == x.shape


def test_exact_pca_oracle_against_sklearn():
    g = np.random.default_rng(1)
    basis = np.linalg.qr(g.standard_normal((16, 16)))[0]
    h = (g.standard_normal((3000, 16)) * np.array([9, 5] + [0.5] * 14)) @ basis + g.standard_normal(16)
    h = h.astype(np.float32)
    mean, comps, evar, z = pca_np.exact_pca(h, 2)
    from sklearn.decomposition import PCA
    sk = PCA(n_components=2, svd_solver="full").fit(h.astype(np.float64))
    np.testing.assert_allclose(mean, sk.mean_, atol=1e-9)
    np.testing.assert_allclose(comps, sk.components_, atol=1e-8)       # same sign convention
    np.testing.assert_allclose(evar, sk.explained_variance_, rtol=1e-9)
    np.testing.assert_allclose(z, sk.transform(h.astype(np.float64)), atol=1e-7)
    # moments form used by the GPU path
    m2, c2, e2 = pca_np.pca_from_moments(len(h), h.astype(np.float64).sum(0), h.astype(np.float64).T @ h.astype(np.float64), 2)
    np.testing.assert_allclose(c2, comps, atol=1e-8)
    # the reference's literal IncrementalPCA spans (nearly) the same plane on a clear spectrum
    ipca, zi = pca_np.incremental_pca_reference(h)
    s = np.linalg.svd(ipca.components_ @ comps.T, compute_uv=False)
    assert s.min() > 0.999


def test_timer_cpu_mirrors_the_reference(capsys):
    """misc/voxel_processing.py:53-70 of the reference: tic / toc(msg), units "ms" and "secs"."""
    import time
    from tomoencoders_b200.misc.voxel_processing import TimerCPU
    t = TimerCPU("ms")
    t.tic()
    time.sleep(0.02)
    ms = t.toc()
    assert 15.0 < ms < 2000.0 and capsys.readouterr().out == ""
    t = TimerCPU("secs")
    t.tic()
    s = t.toc("step")
    assert 0.0 <= s < 1.0 and "TIME: step" in capsys.readouterr().out
