"""Torch-free float64 restatement of the Keras layers the reference's graphs are made of.  TEST INFRASTRUCTURE.

PARITY UNPINNED (no TensorFlow in this image, no golden vectors for network outputs in the reference, SURVEY.md
section 4).  This module exists to remove ONE shared failure mode of the unpinned network oracle
(oracle/nets_torch.py): that torch's ``conv3d`` / ``conv_transpose3d`` / ``max_pool3d`` semantics differ from the
Keras layers they stand in for.  Nothing here calls torch: every layer is written out from the Keras / TensorFlow
definition of the operation as explicit sums over taps (numpy slicing + einsum in float64), NDHWC like Keras.

  Conv3D(f, 3, padding="same")            Unet3D.py:75, autoencoders.py:132
      out[n,z,y,x,o] = b[o] + sum_{a,b,c,i} in[n, z+a-1, y+b-1, x+c-1, i] * K[a,b,c,i,o]       (zeros outside)
      -- TF's conv is a cross-correlation; for an odd kernel and stride 1 "same" pads (k-1)/2 on both sides.
  BatchNormalization (inference)          Unet3D.py:78
      y = gamma * (x - moving_mean) / sqrt(moving_var + eps) + beta
  LeakyReLU(alpha=0.2)                    Unet3D.py:38
  MaxPool3D(p, padding="same")            Unet3D.py:124 (sizes divisible by p: no padding takes part)
  Conv3DTranspose(C, 2, strides=s, padding="same")   Unet3D.py:175
      the gradient of a stride-s VALID/SAME conv with respect to its input: input voxel v contributes
      in[v, i] * K[a,b,c,o,i] to output voxel s*v + (a,b,c) (kernel layout (kz,ky,kx,out,in), no flip); output extent
      s * input extent; for k < s the untouched voxels hold the bias only.
  Conv3D(1, 1) + sigmoid                  Unet3D.py:287
  Dense / Flatten                         autoencoders.py:60-91, 321 (Flatten of NDHWC = C-order reshape)

The graphs (block structure, concat order, weight order) are walked exactly as oracle/nets_torch.py walks them, from
the same flat Keras ``get_weights()`` list.
"""
import numpy as np

LRELU_ALPHA = 0.2
BN_EPS = 1e-5


class _Take:
    def __init__(self, weights):
        self.w, self.i = weights, 0

    def __call__(self, shape):
        a = np.asarray(self.w[self.i], dtype=np.float64)
        self.i += 1
        assert a.shape == tuple(shape), (a.shape, shape)
        return a


def conv3d_same(x, K, b):
    """x (N,D,H,W,Ci), K (k,k,k,Ci,Co), b (Co,) -> (N,D,H,W,Co); k odd."""
    k = K.shape[0]
    r = k // 2
    N, D, H, W, _ = x.shape
    xp = np.zeros((N, D + 2 * r, H + 2 * r, W + 2 * r, x.shape[4]), dtype=np.float64)
    xp[:, r:r + D, r:r + H, r:r + W, :] = x
    out = np.zeros((N, D, H, W, K.shape[4]), dtype=np.float64)
    for a in range(k):
        for bb in range(k):
            for c in range(k):
                out += np.einsum("nzyxi,io->nzyxo", xp[:, a:a + D, bb:bb + H, c:c + W, :], K[a, bb, c])
    return out + b


def batchnorm(x, gamma, beta, mean, var):
    return gamma * (x - mean) / np.sqrt(var + BN_EPS) + beta


def lrelu(x):
    return np.where(x > 0, x, LRELU_ALPHA * x)


def maxpool(x, p):
    N, D, H, W, C = x.shape
    assert D % p == 0 and H % p == 0 and W % p == 0
    return x.reshape(N, D // p, p, H // p, p, W // p, p, C).max(axis=(2, 4, 6))


def conv3d_transpose(x, K, b, s):
    """x (N,D,H,W,Ci), K (k,k,k,Co,Ci), stride s >= k -> (N,sD,sH,sW,Co)."""
    k = K.shape[0]
    assert k <= s
    N, D, H, W, _ = x.shape
    out = np.zeros((N, D * s, H * s, W * s, K.shape[3]), dtype=np.float64)
    for a in range(k):
        for bb in range(k):
            for c in range(k):
                out[:, a::s, bb::s, c::s, :] += np.einsum("nzyxi,oi->nzyxo", x, K[a, bb, c])
    return out + b


def _conv_block(x, take, cout, bn, training=False):
    cin = x.shape[-1]
    y = conv3d_same(x, take((3, 3, 3, cin, cout)), take((cout,)))
    if bn:
        gamma, beta, mean, var = take((cout,)), take((cout,)), take((cout,)), take((cout,))
        if training:        # Keras BatchNormalization(training=True): batch mean and BIASED batch variance over (N, D, H, W)
            mean, var = y.mean(axis=(0, 1, 2, 3)), y.var(axis=(0, 1, 2, 3))
        y = batchnorm(y, gamma, beta, mean, var)
    return lrelu(y)


def unet_forward(x, weights, n_filters=(16, 32, 64), n_blocks=3, activation="lrelu", batch_norm=True, kern_size=3,
                 kern_size_upconv=2, isconcat=None, pool_size=2, return_logits=False, training=False):
    """build_Unet_3D (Unet3D.py:199-291).  x (N,D,H,W,1) -> (N,D,H,W,1) float64 probabilities or logits.
    training=True normalises every BatchNormalization layer with the statistics of the batch (model.fit)."""
    assert activation == "lrelu" and kern_size == 3 and kern_size_upconv == 2
    if isconcat is None:
        isconcat = [False] * n_blocks
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    take = _Take(weights)
    code = np.asarray(x, dtype=np.float64)
    skips = []
    for i in range(n_blocks):
        code = _conv_block(code, take, n_filters[i], batch_norm, training)
        code = _conv_block(code, take, 2 * n_filters[i], batch_norm, training)
        skips.append(code)
        if pool_size[i] > 1:
            code = maxpool(code, pool_size[i])
    nf = code.shape[-1]
    code = _conv_block(code, take, nf, batch_norm, training)
    dec = _conv_block(code, take, 2 * nf, batch_norm, training)
    for i in range(n_blocks - 1, -1, -1):
        if pool_size[i] > 1:
            c = dec.shape[-1]
            dec = lrelu(conv3d_transpose(dec, take((2, 2, 2, c, c)), take((c,)), pool_size[i]))
            if isconcat[i]:
                dec = np.concatenate([dec, skips[i]], axis=-1)          # up-sampled first, skip second (Unet3D.py:180)
        dec = _conv_block(dec, take, 2 * n_filters[i], batch_norm, training)
        dec = _conv_block(dec, take, 2 * n_filters[i], batch_norm, training)
    cin = dec.shape[-1]
    logits = np.einsum("nzyxi,io->nzyxo", dec, take((1, 1, 1, cin, 1))[0, 0, 0]) + take((1,))
    assert take.i == len(weights)
    return logits if return_logits else 1.0 / (1.0 + np.exp(-logits))


def keras_bce(p, y):
    """tf.keras.losses.BinaryCrossentropy on probabilities (surface_segmenter.py:322-323; Keras backend: clip to
    [1e-7, 1 - 1e-7], log(p + 1e-7), mean over all voxels)."""
    eps = 1e-7
    pc = np.clip(p, eps, 1.0 - eps)
    return float(-(y * np.log(pc + eps) + (1.0 - y) * np.log(1.0 - pc + eps)).mean())


def unet_training_loss(x, y, weights, **model_params):
    """Loss of one fit step of the segmenter (BatchNormalization in training mode + BCE), float64, no autograd: the
    finite-difference reference for the gradient oracle (oracle/train_torch.py)."""
    p = unet_forward(x, weights, training=True, **model_params)
    return keras_bce(p, np.asarray(y, dtype=np.float64))


def _dense(x, take, nout, bn):
    y = x @ take((x.shape[1], nout)) + take((nout,))
    if bn:
        y = batchnorm(y, take((nout,)), take((nout,)), take((nout,)), take((nout,)))
    return lrelu(y)


def encoder_forward(x, weights, n_filters=(32, 64), n_blocks=2, activation="lrelu", batch_norm=True, kern_size=3,
                    kern_size_upconv=2, hidden_units=(128, 32, 2), pool_size=2, POOL_FLAG=True):
    """build_encoder_r (autoencoders.py:246-347).  x (N,D,H,W,1) -> (N, latent) float64."""
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    take = _Take(weights)
    code = np.asarray(x, dtype=np.float64)
    for i in range(n_blocks):
        code = _conv_block(code, take, n_filters[i], batch_norm)
        code = _conv_block(code, take, 2 * n_filters[i], batch_norm)
        code = maxpool(code, pool_size[i])
    if POOL_FLAG:
        code = maxpool(code, 2)
    code = code.reshape(code.shape[0], -1)
    for n_hidden in list(hidden_units)[:-1]:
        code = _dense(code, take, n_hidden, batch_norm)
    z = _dense(code, take, hidden_units[-1], True)
    assert take.i == len(weights)
    return z


def decoder_forward(z, weights, preflatten_shape, n_filters=(32, 64), n_blocks=2, activation="lrelu", batch_norm=True,
                    kern_size=3, kern_size_upconv=2, hidden_units=(128, 32, 2), pool_size=2, POOL_FLAG=True):
    """build_decoder_r (autoencoders.py:350-447).  z (N, latent) -> (N,D,H,W,1) float64 probabilities."""
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    take = _Take(weights)
    dec = np.asarray(z, dtype=np.float64)
    for n_hidden in list(hidden_units)[::-1][1:]:
        dec = _dense(dec, take, n_hidden, batch_norm)
    dec = _dense(dec, take, int(np.prod(preflatten_shape)), batch_norm)
    dec = dec.reshape((dec.shape[0],) + tuple(preflatten_shape))
    if POOL_FLAG:
        c = dec.shape[-1]
        dec = lrelu(conv3d_transpose(dec, take((2, 2, 2, c, c)), take((c,)), 2))
    for i in range(n_blocks - 1, -1, -1):
        c = dec.shape[-1]
        dec = lrelu(conv3d_transpose(dec, take((2, 2, 2, c, c)), take((c,)), pool_size[i]))
        dec = _conv_block(dec, take, n_filters[i], batch_norm)
        dec = _conv_block(dec, take, n_filters[i], batch_norm)
    cin = dec.shape[-1]
    logits = np.einsum("nzyxi,io->nzyxo", dec, take((1, 1, 1, cin, 1))[0, 0, 0]) + take((1,))
    assert take.i == len(weights)
    return 1.0 / (1.0 + np.exp(-logits))
