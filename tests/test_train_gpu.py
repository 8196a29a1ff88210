"""GPU: one training step of the U-net segmenter against the torch-CPU
oracle (oracle/train_torch.py): loss, every parameter gradient, the Adam update and the BN moving
statistics.  fp32 check mode: tight tolerances; bf16 tensor-core mode: loss within 1e-2 relative and
gradient cosine similarity per tensor."""
import numpy as np
import pytest
import torch

from oracle import nets_torch as O, train_torch as T, weights as OW
from tomoencoders_b200 import Patches
from tomoencoders_b200.neural_nets._train import CAETrainer, UnetTrainer

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
    # 60 fit steps: with 30 the accuracy below depends on the draw of the initial weights (0.85 .. 0.99 observed); with 60 every
    # seed tried reaches > 0.999 (tools/scratch/train_flaky.py)
    hist = fe.train([X], [Y], (32, 32, 32), 3, 8, 0.1, True, 0.05, 6, steps_per_epoch=10, verbose=0)
    assert hist[-1] < hist[0] and hist[-1] < 0.5, hist
    w1 = fe.models["segmenter"].get_weights()
    assert any(np.abs(a - b).max() > 1e-4 for a, b in zip(w0, w1))
    p = Patches(X.shape, "regular-grid", patch_size=(32,) * 3)
    yp = fe.predict_patches("segmenter", p.extract(X, (32,) * 3)[..., np.newaxis], 8, None, min_max=(0.0, 1.0))
    acc = float((yp[..., 0] == p.extract(Y, (32,) * 3)).mean())
    assert acc > 0.9, acc


CAE = dict(n_filters=[16, 32], n_blocks=2, activation="lrelu", batch_norm=True, pool_size=[2, 2], hidden_units=[64, 16],
           POOL_FLAG=True)


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_autoencoder_step_matches_oracle(mode):
    """RegularizedAutoencoder.train_step (autoencoders.py:498-526): MSE + KL / 250 through encoder and decoder."""
    size = (32, 32, 32)
    se, _, _ = O.encoder_spec(size, **CAE)
    sd = O.decoder_spec(size, **CAE)
    we, wd = OW.draw(se, 31), OW.draw(sd, 32)
    nb = 4 if mode == "fp32" else 16          # BatchNormalization of the dense layers runs over the batch: see below
    x = np.random.default_rng(9).uniform(0, 1, (nb,) + size + (1,)).astype(np.float32)
    tot, mse, kl, grads_ref, moving_ref = T.cae_loss_and_grads(x, we, wd, size, **CAE)
    tr = CAETrainer(CAE, size, nb, mode=mode, weights=we + wd)
    loss = float(tr.forward_backward(x).item())
    l3 = tr.losses()
    tol = 1e-5 if mode == "fp32" else 1e-2
    assert abs(loss - tot) < tol * abs(tot) and abs(l3[1] - mse) < tol * abs(mse) and abs(l3[2] - kl) < 5 * tol * abs(kl)
    grads = tr.get_gradients()
    cos = {}
    gmax = max(np.linalg.norm(g) for g in grads_ref if g is not None)
    for i, (kind, _) in enumerate(tr.spec):
        if grads_ref[i] is None or (kind in ("conv_b", "dense_b") and i != len(tr.spec) - 1):
            continue                                   # biases in front of a BN: zero gradient, pure noise
        if np.linalg.norm(grads_ref[i]) < 1e-9 * gmax:
            assert np.linalg.norm(grads[i]) < 1e-4 * gmax   # mathematically zero (here: the beta in front of the two pools)
            continue
        cos[(i, kind)] = _cos(grads[i], grads_ref[i])
    worst = sorted(cos.items(), key=lambda kv: kv[1])[:3]
    if mode == "fp32":
        assert min(cos.values()) > 0.999, worst
    else:
        dec = {k: v for k, v in cos.items() if k[0] >= tr.n_enc}
        enc = {k: v for k, v in cos.items() if k[0] < tr.n_enc}
        print("bf16 autoencoder gradient cosines: decoder min %.4f median %.4f | encoder min %.4f median %.4f" %
              (min(dec.values()), float(np.median(list(dec.values()))), min(enc.values()), float(np.median(list(enc.values())))))
        # (with a batch of 4 the BatchNormalization of the dense layers -- statistics over 4 samples -- amplifies every
        # bf16 rounding and the encoder-side cosines drop to ~0.45; at 16 samples the comparison is well conditioned)
        assert min(dec.values()) > 0.9 and float(np.median(list(dec.values()))) > 0.98, worst
        assert min(enc.values()) > 0.85 and float(np.median(list(enc.values()))) > 0.93, worst
    tr.apply_gradients()
    w_new = tr.get_weights()
    for i, (kind, _) in enumerate(tr.spec):
        if kind in ("bn_mean", "bn_var"):
            np.testing.assert_allclose(w_new[i], moving_ref[i], rtol=5e-2 if mode == "bf16" else 1e-4, atol=5e-3 if mode == "bf16" else 1e-5)


def test_selfsupervised_cae_train_api_lowers_the_loss():
    from scipy.ndimage import gaussian_filter
    from tomoencoders_b200.neural_nets import SelfSupervisedCAE
    g = np.random.default_rng(2)
    b = gaussian_filter(g.standard_normal((80, 80, 80)), 3)
    X = ((b - b.min()) / (b.max() - b.min())).astype(np.float32)
    np.random.seed(1)
    ae = SelfSupervisedCAE(model_initialization="define-new", model_size=(32, 32, 32), descriptor_tag="t", mode="bf16", **CAE)
    hist = ae.train([X], batch_size=8, sampling_method="random-fixed-width", n_epochs=3, random_rotate=True, steps_per_epoch=15,
                    verbose=0)
    assert hist[-1] < hist[0], hist
    x = Patches(X.shape, "regular-grid", patch_size=(32,) * 3).extract(X, (32,) * 3)[..., np.newaxis]
    z = ae.predict_embeddings(x, 8, min_max=None)
    assert z.shape == (len(x), 16) and np.isfinite(z).all()
