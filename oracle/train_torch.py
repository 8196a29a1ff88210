"""torch-CPU restatement of one Keras training step of the U-net segmenter.  TEST INFRASTRUCTURE.

PARITY UNPINNED (no TensorFlow here, no golden vectors in the reference).  Follows
  SurfaceSegmenter.train / _build_models   /root/reference/tomo_encoders/neural_nets/surface_segmenter.py:43-66, 297-324
  build_Unet_3D                            neural_nets/Unet3D.py:199-291 (graph shared with oracle/nets_torch.py)
and the Keras conventions of SURVEY.md appendix A.3: BatchNormalization(momentum=0.9, epsilon=1e-5) in
training mode (batch mean, biased batch variance; moving statistics <- 0.9 moving + 0.1 batch),
BinaryCrossentropy on probabilities (Keras backend: clip to [1e-7, 1 - 1e-7], log(p + 1e-7), mean over
all voxels), Adam() defaults (lr 1e-3, beta 0.9 / 0.999, eps 1e-7 outside the square root, bias
correction folded into the step size).
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import nets_torch as O

BN_MOMENTUM = 0.9
TRAINABLE = ("conv_k", "conv_b", "convT_k", "convT_b", "dense_k", "dense_b", "bn_gamma", "bn_beta")


class _WT(O._W):
    """Weight cursor for training: trainable arrays become autograd leaves; BN layers normalise
    with batch statistics and record the updated moving statistics."""

    def __init__(self, weights, spec, dtype):
        super().__init__(weights, dtype)
        self.kinds = [k for k, _ in spec]
        self.leaves = {}
        self.new_moving = {}
        self.training = True

    def take(self, kind, shape):
        idx = self.i
        t = super().take(kind, shape)
        if kind in TRAINABLE:
            t = t.clone().requires_grad_(True)
            self.leaves[idx] = t
        else:
            self.leaves[idx] = t
        return t


def _bn_train(t, W, c):
    g_i = W.i
    g = W.take("bn_gamma", (c,))
    b = W.take("bn_beta", (c,))
    m = W.take("bn_mean", (c,))
    v = W.take("bn_var", (c,))
    dims = (0,) + tuple(range(2, t.dim()))
    mean = t.mean(dims)
    var = t.var(dims, unbiased=False)
    shape = (1, c) + (1,) * (t.dim() - 2)
    out = g.view(shape) * (t - mean.view(shape)) / torch.sqrt(var.view(shape) + O.BN_EPS) + b.view(shape)
    W.new_moving[g_i + 2] = (BN_MOMENTUM * m + (1 - BN_MOMENTUM) * mean.detach())
    W.new_moving[g_i + 3] = (BN_MOMENTUM * v + (1 - BN_MOMENTUM) * var.detach())
    return out


def keras_bce(p, y):
    eps = 1e-7
    pc = torch.clamp(p, eps, 1.0 - eps)
    return -(y * torch.log(pc + eps) + (1.0 - y) * torch.log(1.0 - pc + eps)).mean()


def unet_loss_and_grads(x, y, weights, dtype=torch.float64, **model_params):
    """x (N,D,H,W,1) float, y (N,D,H,W,1) {0,1}.  Returns (loss, grads, new_moving) with grads a list
    in Keras get_weights() order (None for the moving statistics) and new_moving {index: array}."""
    spec = O.unet_spec(**model_params)
    W = _WT(weights, spec, dtype)
    old_bn = O._bn
    O._bn = _bn_train
    try:
        t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype).permute(0, 4, 1, 2, 3)
        logits = O.unet_graph(t, W, return_logits=True, **model_params)
    finally:
        O._bn = old_bn
    yt = torch.from_numpy(np.ascontiguousarray(y)).to(dtype).permute(0, 4, 1, 2, 3)
    loss = keras_bce(torch.sigmoid(logits), yt)
    loss.backward()
    grads = []
    for i, kind in enumerate(W.kinds):
        leaf = W.leaves[i]
        grads.append(leaf.grad.detach().numpy().astype(np.float64) if kind in TRAINABLE else None)
    new_moving = {i: v.numpy().astype(np.float64) for i, v in W.new_moving.items()}
    return float(loss.item()), grads, new_moving


class KerasAdam:
    """tf.keras.optimizers.Adam() defaults."""

    def __init__(self, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        self.lr, self.b1, self.b2, self.eps = lr, beta_1, beta_2, epsilon
        self.t = 0
        self.m = self.v = None

    def step(self, weights, grads):
        if self.m is None:
            self.m = [None if g is None else np.zeros_like(g) for g in grads]
            self.v = [None if g is None else np.zeros_like(g) for g in grads]
        self.t += 1
        lr_t = self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t)
        out = []
        for w, g, i in zip(weights, grads, range(len(weights))):
            if g is None:
                out.append(np.asarray(w))
                continue
            self.m[i] = self.b1 * self.m[i] + (1 - self.b1) * g
            self.v[i] = self.b2 * self.v[i] + (1 - self.b2) * g * g
            out.append(np.asarray(w, dtype=np.float64) - lr_t * self.m[i] / (np.sqrt(self.v[i]) + self.eps))
        return out


def unet_train_step(x, y, weights, opt, dtype=torch.float64, **model_params):
    """One fit step: returns (loss, new_weights) with the moving statistics updated."""
    loss, grads, new_moving = unet_loss_and_grads(x, y, weights, dtype, **model_params)
    new_w = opt.step(weights, grads)
    for i, v in new_moving.items():
        new_w[i] = v
    return loss, [np.asarray(w, dtype=np.float32) for w in new_w], grads


def keras_mse_kl(x, p, weight):
    """RegularizedAutoencoder.train_step (autoencoders.py:498-526): mean(mse(x, p)) + weight * mean(kl_divergence(x, p)),
    Keras' kl_divergence clips both arguments to [1e-7, 1] and sums y_true * log(y_true / y_pred) over the channel axis."""
    eps = 1e-7
    mse = ((x - p) ** 2).mean()
    yt, yp = torch.clamp(x, eps, 1.0), torch.clamp(p, eps, 1.0)
    kl = (yt * torch.log(yt / yp)).mean()
    return mse + weight * kl, mse, kl


def cae_loss_and_grads(x, weights_enc, weights_dec, model_size, weight=1.0 / 250.0, dtype=torch.float64, **model_params):
    """x (N,D,H,W,1) in [0,1].  Returns (total, mse, kl, grads_enc + grads_dec (Keras order; None for moving statistics),
    new_moving {index in the concatenated list: array})."""
    spec_e, flatten_shape, preflatten = O.encoder_spec(model_size, **model_params)
    spec_d = O.decoder_spec(model_size, **model_params)
    W = _WT(list(weights_enc) + list(weights_dec), list(spec_e) + list(spec_d), dtype)
    old_bn = O._bn
    O._bn = _bn_train
    try:
        t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype).permute(0, 4, 1, 2, 3)
        z, _, _ = O.encoder_graph(t, W, **model_params)
        assert W.i == len(spec_e)
        p = O.decoder_graph(z, W, flatten_shape, preflatten, **model_params)
    finally:
        O._bn = old_bn
    total, mse, kl = keras_mse_kl(t, p, weight)
    total.backward()
    grads = []
    for i, kind in enumerate(W.kinds):
        leaf = W.leaves[i]
        grads.append(leaf.grad.detach().numpy().astype(np.float64) if kind in TRAINABLE and leaf.grad is not None else None)
    new_moving = {i: v.numpy().astype(np.float64) for i, v in W.new_moving.items()}
    return float(total.item()), float(mse.item()), float(kl.item()), grads, new_moving
