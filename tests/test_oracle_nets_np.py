"""CPU: the torch-free float64 restatement of the Keras layers (oracle/nets_np.py, explicit sums over taps) against the
torch-CPU oracle (oracle/nets_torch.py) in float64 -- two independent restatements of the same Keras graphs must agree to
rounding, so "torch semantics != Keras semantics" is not a shared failure mode of every network parity number."""
import numpy as np
import torch

from oracle import nets_np as NP, nets_torch as O, weights as OW


def test_unet_np_matches_torch_oracle_fp64():
    for params, size in ((dict(n_filters=[4, 6], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True, False],
                               pool_size=[2, 2]), (8, 12, 16)),
                         (dict(n_filters=[3, 5], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True, True],
                               pool_size=[2, 4]), (8, 8, 16)),          # k = 2 < stride 4: bias-only gaps
                         (dict(n_filters=[4], n_blocks=1, activation="lrelu", batch_norm=False, isconcat=[True],
                               pool_size=[1]), (6, 5, 7))):             # pool 1: no pool, no upconv (Unet3D.py:120,175)
        w = OW.draw(O.unet_spec(**params), 3)
        x = np.random.default_rng(1).uniform(0, 1, (2,) + size + (1,))
        a = NP.unet_forward(x, w, return_logits=True, **params)
        b = O.unet_forward(x, w, dtype=torch.float64, return_logits=True, **params)
        assert a.shape == b.shape
        assert np.abs(a - b).max() < 1e-10 * max(1.0, np.abs(b).max())
        p = NP.unet_forward(x, w, **params)
        assert np.abs(p - O.unet_forward(x, w, dtype=torch.float64, **params)).max() < 1e-12


def test_encoder_decoder_np_match_torch_oracle_fp64():
    params = dict(n_filters=[4, 6], n_blocks=2, activation="lrelu", batch_norm=True, pool_size=[2, 2],
                  hidden_units=[16, 5], POOL_FLAG=True)
    size = (16, 16, 16)
    spec, flat, pre = O.encoder_spec(size, **params)
    we = OW.draw(spec, 4)
    wd = OW.draw(O.decoder_spec(size, **params), 5)
    x = np.random.default_rng(2).uniform(0, 1, (3,) + size + (1,))
    z = NP.encoder_forward(x, we, **params)
    zt = O.encoder_forward(x, we, dtype=torch.float64, **params)
    assert np.abs(z - zt).max() < 1e-10
    y = NP.decoder_forward(z, wd, pre, **params)
    yt = O.decoder_forward(zt, wd, size, dtype=torch.float64, **params)
    assert y.shape == yt.shape == (3,) + size + (1,)
    assert np.abs(y - yt).max() < 1e-10


def test_single_layer_known_answers():
    """Hand-checkable cases of the layer definitions themselves."""
    x = np.zeros((1, 3, 3, 3, 1))
    x[0, 1, 1, 1, 0] = 1.0
    K = np.arange(27, dtype=np.float64).reshape(3, 3, 3, 1, 1)
    y = NP.conv3d_same(x, K, np.zeros(1))
    # cross-correlation: the impulse at the centre shows the kernel FLIPPED around it
    np.testing.assert_array_equal(y[0, ..., 0], K[::-1, ::-1, ::-1, 0, 0])
    # corner output sees only the 8 in-bounds taps (zero padding)
    ones = np.ones((1, 3, 3, 3, 1))
    assert NP.conv3d_same(ones, np.ones((3, 3, 3, 1, 1)), np.zeros(1))[0, 0, 0, 0, 0] == 8.0
    # transposed conv: voxel v writes K[a,b,c] at 2 v + (a,b,c)
    xt = np.zeros((1, 2, 2, 2, 1))
    xt[0, 1, 0, 1, 0] = 2.0
    Kt = np.arange(8, dtype=np.float64).reshape(2, 2, 2, 1, 1)
    yt = NP.conv3d_transpose(xt, Kt, np.array([0.5]), 2)
    exp = np.full((4, 4, 4), 0.5)
    exp[2:4, 0:2, 2:4] += 2.0 * Kt[..., 0, 0]
    np.testing.assert_array_equal(yt[0, ..., 0], exp)
    # stride 4, kernel 2: bias-only gaps
    y4 = NP.conv3d_transpose(xt, Kt, np.array([0.5]), 4)
    assert y4.shape == (1, 8, 8, 8, 1) and y4[0, 6, 2, 6, 0] == 0.5 and y4[0, 4, 0, 4, 0] == 0.5
    assert y4[0, 5, 1, 5, 0] == 0.5 + 2.0 * 7


def test_training_oracle_loss_and_gradients_against_torch_free_finite_differences():
    """The gradient oracle of the training step (oracle/train_torch.py: torch autograd through BatchNormalization in
    training mode + Keras BCE) against a torch-free pin: the numpy float64 loss of the same step equals its loss, and
    central finite differences of that numpy loss along random directions of every trainable array equal the
    autograd gradient's directional derivative."""
    from oracle import train_torch as TT
    P = dict(n_filters=[2, 3], n_blocks=2, activation="lrelu", batch_norm=True, isconcat=[True, True], pool_size=[2, 2])
    spec = O.unet_spec(**P)
    w = [np.asarray(a, dtype=np.float64) for a in OW.draw(spec, 5)]
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 1, (3, 8, 8, 8, 1))
    y = (rng.uniform(0, 1, (3, 8, 8, 8, 1)) < 0.4).astype(np.float64)
    loss_t, grads, _ = TT.unet_loss_and_grads(x, y, w, **P)
    loss_n = NP.unet_training_loss(x, y, w, **P)
    assert abs(loss_n - loss_t) < 1e-12 * max(1.0, abs(loss_t))
    checked = 0
    for i, (kind, shape) in enumerate(spec):
        if grads[i] is None:
            continue                                            # moving statistics: not trainable, not used in training mode
        d = rng.standard_normal(shape)
        d /= np.linalg.norm(d)
        h = 1e-5
        wp, wm = list(w), list(w)
        wp[i], wm[i] = w[i] + h * d, w[i] - h * d
        fd = (NP.unet_training_loss(x, y, wp, **P) - NP.unet_training_loss(x, y, wm, **P)) / (2 * h)
        an = float((grads[i] * d).sum())
        assert abs(fd - an) < 1e-6 * max(1e-3, abs(an)) + 1e-9, (i, kind, fd, an)
        checked += 1
    assert checked >= 30


def test_toeplitz_formulation_of_the_first_layer_equals_the_direct_convolution():
    """The index algebra of csrc/conv1_tz.cu (aligned 8-wide planes, two x phases, banded weight tiles) restated in numpy
    equals Conv3D(f, 3, 'same') on one channel, exactly up to float64 rounding, for every width the kernel takes."""
    from oracle import conv1_toeplitz_np as TZ
    rng = np.random.default_rng(4)
    for shape, co in (((2, 4, 5, 8), 16), ((1, 3, 4, 24), 16), ((1, 2, 3, 64), 5)):
        x = rng.standard_normal(shape)
        w = rng.standard_normal((3, 3, 3, 1, co))
        b = rng.standard_normal(co)
        want = NP.conv3d_same(x[..., None], w, b)
        got = TZ.conv1_toeplitz(x, w, b)
        assert np.abs(got - want).max() < 1e-12 * np.abs(want).max()
    # the banded tiles hold each weight exactly XT times per (phase, tap) and nothing else
    w = rng.standard_normal((3, 3, 3, 1, 16))
    B = TZ.banded_weights(w)
    assert B.shape == (2, 9, 16, 64) and np.count_nonzero(B) == 2 * 9 * 4 * 3 * 16
    assert not B[0, :, :7].any() and not B[0, :, 13:].any() and not B[1, :, :3].any() and not B[1, :, 9:].any()
