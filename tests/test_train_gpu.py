"""GPU: one training step of the U-net segmenter against the torch-CPU
oracle (oracle/train_torch.py): loss, every parameter gradient, the Adam update and the BN moving
statistics.  fp32 check mode: tight tolerances; bf16 tensor-core mode: loss within 1e-2 relative and
gradient cosine similarity per tensor."""
import numpy as np
import pytest
import torch

from oracle import nets_torch as O, train_torch as T, weights as OW
from tomoencoders_b200 import Patches
from tomoencoders_b200.neural_nets._train import UnetTrainer

pytestmark = pytest.mark.gpu

SMALL = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True, True], pool_size=[2, 2])


def _data(n, size, seed):
    g = np.random.default_rng(seed)
    x = g.uniform(0, 1, (n,) + size + (1,)).astype(np.float32)
    y = (g.uniform(0, 1, (n,) + size + (1,)) < 0.4).astype(np.uint8)
    return x, y


def _cos(a, b):
    a, b = np.asarray(a, np.float64).ravel(), np.asarray(b, np.float64).ravel()
    return float(a @ b / (np.linalg.norm(a) * np.linalg.norm(b) + 1e-300))


@pytest.mark.parametrize("params,size", [(SMALL, (16, 16, 16)),
                                         (dict(SMALL, isconcat=[False, True], pool_size=[2, 2]), (8, 16, 24))])
def test_fp32_step_matches_oracle(params, size):
    w = OW.draw(O.unet_spec(**params), 21)
    x, y = _data(2, size, 5)
    loss_ref, grads_ref, moving_ref = T.unet_loss_and_grads(x, y, w, **params)
    tr = UnetTrainer(params, size, 2, mode="fp32", weights=w)
    loss = float(tr.forward_backward(x, y).item())
    assert abs(loss - loss_ref) < 1e-5 * max(1.0, abs(loss_ref))
    grads = tr.get_gradients()
    for i, (kind, _) in enumerate(tr.spec):
        if grads_ref[i] is None:
            continue
        scale = np.abs(grads_ref[i]).max() + 1e-12
        err = np.abs(grads[i] - grads_ref[i]).max() / scale
        # conv biases feeding a BatchNormalization have an exactly zero gradient (BN removes the mean)
        if kind == "conv_b" and scale < 1e-9:
            assert np.abs(grads[i]).max() < 1e-6
            continue
        assert err < 1e-2 and _cos(grads[i], grads_ref[i]) > 0.9999, (i, kind, err, _cos(grads[i], grads_ref[i]))
    # Adam + moving statistics.  The first Adam step moves every weight by lr * g / (|g| + 1e-7): with
    # gradients of a mean loss (|g| ~ 1e-6) that is extremely sensitive to g, so the update is
    # checked against Keras' rule applied to OUR gradients (the gradients themselves were compared
    # above), and the moving statistics against the oracle's.
    opt = T.KerasAdam()
    g64 = [None if g is None else np.asarray(g, np.float64) for g in grads]
    for i, (kind, _) in enumerate(tr.spec):
        if grads_ref[i] is None:
            g64[i] = None
    w_exp = opt.step([np.asarray(a, np.float64) for a in w], g64)
    tr.apply_gradients()
    w_new = tr.get_weights()
    for i, (kind, _) in enumerate(tr.spec):
        if kind in ("bn_mean", "bn_var"):
            np.testing.assert_allclose(w_new[i], moving_ref[i], rtol=1e-4, atol=1e-5)
        else:
            assert np.abs(w_new[i] - w_exp[i]).max() < 1e-6, (i, kind)


def test_bf16_step_is_close_to_oracle():
    params, size = SMALL, (32, 32, 32)
    w = OW.draw(O.unet_spec(**params), 22)
    x, y = _data(2, size, 6)
    loss_ref, grads_ref, _ = T.unet_loss_and_grads(x, y, w, dtype=torch.float32, **params)
    tr = UnetTrainer(params, size, 2, mode="bf16", weights=w)
    loss = float(tr.forward_backward(x, y).item())
    assert abs(loss - loss_ref) < 1e-2 * abs(loss_ref)
    grads = tr.get_gradients()
    cosines = {}
    for i, (kind, _) in enumerate(tr.spec):
        if grads_ref[i] is None or (kind == "conv_b" and i != len(tr.spec) - 1):
            continue                                    # conv biases in front of a BN: zero gradient, pure noise
        cosines[(i, kind)] = _cos(grads[i], grads_ref[i])
    print("bf16 gradient cosines: min %.4f median %.4f" % (min(cosines.values()), float(np.median(list(cosines.values())))))
    # 8-bit mantissas through ~20 layers of forward and backward: the deepest (first-layer) gradients
    # are the least accurate
    assert min(cosines.values()) > 0.85, sorted(cosines.items(), key=lambda kv: kv[1])[:3]
    assert min(v for (i, k), v in cosines.items() if k in ("conv_k", "convT_k")) > 0.95
    assert float(np.median(list(cosines.values()))) > 0.97


def test_loss_decreases_over_a_few_steps():
    params, size = SMALL, (32, 32, 32)
    g = np.random.default_rng(7)
    x = g.uniform(0, 1, (4,) + size + (1,)).astype(np.float32)
    y = (x > 0.5).astype(np.uint8)                      # learnable target
    tr = UnetTrainer(params, size, 4, mode="bf16", seed=3)
    losses = [tr.train_step(x, y) for _ in range(12)]
    assert losses[-1] < 0.8 * losses[0], losses


def test_surface_segmenter_train_api_learns_a_synthetic_volume():
    """SurfaceSegmenter.train (surface_segmenter.py:43-66): random multi-scale patches with blank
    rejection, noise and rot90 augmentation; the loss must fall and the trained weights must be the
    ones predict_patches uses afterwards."""
    from scipy.ndimage import gaussian_filter
    from tomoencoders_b200.neural_nets import SurfaceSegmenter
    g = np.random.default_rng(1)
    b = gaussian_filter(g.standard_normal((96, 96, 96)), 3)
    Y = (b > np.quantile(b, 0.6)).astype(np.uint8)
    X = (Y + 0.1 * g.standard_normal(Y.shape)).astype(np.float32)
    np.random.seed(0)
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="t", mode="bf16", **SMALL)
    w0 = fe.models["segmenter"].get_weights()
    hist = fe.train([X], [Y], (32, 32, 32), 3, 8, 0.1, True, 0.05, 3, steps_per_epoch=10, verbose=0)
    assert hist[-1] < hist[0] and hist[-1] < 0.5, hist
    w1 = fe.models["segmenter"].get_weights()
    assert any(np.abs(a - b).max() > 1e-4 for a, b in zip(w0, w1))
    p = Patches(X.shape, "regular-grid", patch_size=(32,) * 3)
    yp = fe.predict_patches("segmenter", p.extract(X, (32,) * 3)[..., np.newaxis], 8, None, min_max=(0.0, 1.0))
    acc = float((yp[..., 0] == p.extract(Y, (32,) * 3)).mean())
    assert acc > 0.9, acc
