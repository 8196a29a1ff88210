"""Weight specifications (Keras ``get_weights()`` order, layouts and shapes) of the graphs the
reference builds, and Keras-default initialisation for ``model_initialization="define-new"``.

Mirrors the layer creation order of build_Unet_3D (/root/reference/tomo_encoders/neural_nets/
Unet3D.py:199-291), build_encoder_r (neural_nets/autoencoders.py:246-347) and build_decoder_r
(:350-447); the planner in csrc/net.cu consumes the flat blob in exactly this order and the two are
cross-checked through te_net_weight_count().

Entry = (kind, shape), kind in conv_k, conv_b, convT_k, convT_b, dense_k, dense_b, bn_gamma,
bn_beta, bn_mean, bn_var.  Conv3D kernel (kz,ky,kx,Cin,Cout); Conv3DTranspose kernel
(kz,ky,kx,Cout,Cin); Dense kernel (in,out).
"""
import numpy as np


def _pools(pool_size, n_blocks):
    if isinstance(pool_size, (int, np.integer)):
        return [int(pool_size)] * n_blocks
    if len(pool_size) != n_blocks:
        raise ValueError("list length must be equal to number of blocks")
    return [int(p) for p in pool_size]


def _conv(spec, cin, cout, k, bn):
    spec.append(("conv_k", (k, k, k, cin, cout)))
    spec.append(("conv_b", (cout,)))
    if bn:
        _bn(spec, cout)
    return cout


def _bn(spec, c):
    for kind in ("bn_gamma", "bn_beta", "bn_mean", "bn_var"):
        spec.append((kind, (c,)))


def _upconv(spec, c, k):
    spec.append(("convT_k", (k, k, k, c, c)))
    spec.append(("convT_b", (c,)))
    return c


def _dense(spec, nin, nout, bn):
    spec.append(("dense_k", (nin, nout)))
    spec.append(("dense_b", (nout,)))
    if bn:
        _bn(spec, nout)
    return nout


def unet_spec(n_filters=(16, 32, 64), n_blocks=3, activation="lrelu", batch_norm=True, kern_size=3,
              kern_size_upconv=2, isconcat=None, pool_size=2):
    pools = _pools(pool_size, n_blocks)
    isconcat = [False] * n_blocks if isconcat is None else list(isconcat)
    spec, c, skips = [], 1, []
    for i in range(n_blocks):
        c = _conv(spec, c, n_filters[i], kern_size, batch_norm)
        c = _conv(spec, c, 2 * n_filters[i], kern_size, batch_norm)
        skips.append(c)
    nf = c
    c = _conv(spec, c, nf, kern_size, batch_norm)
    c = _conv(spec, c, 2 * nf, kern_size, batch_norm)
    for i in range(n_blocks - 1, -1, -1):
        if pools[i] > 1:
            c = _upconv(spec, c, kern_size_upconv)
            if isconcat[i]:
                c = c + skips[i]
        c = _conv(spec, c, 2 * n_filters[i], kern_size, batch_norm)
        c = _conv(spec, c, 2 * n_filters[i], kern_size, batch_norm)
    spec.append(("conv_k", (1, 1, 1, c, 1)))
    spec.append(("conv_b", (1,)))
    return spec


def encoder_shapes(model_size, n_filters, n_blocks, pool_size, POOL_FLAG):
    pools = _pools(pool_size, n_blocks)
    size = [int(s) for s in model_size]
    for p in pools:
        if any(s % p for s in size):
            raise ValueError("pool size must divide the spatial extent")
        size = [s // p for s in size]
    if POOL_FLAG:
        if any(s % 2 for s in size):
            raise ValueError("pool size must divide the spatial extent")
        size = [s // 2 for s in size]
    preflatten = tuple(size) + (2 * n_filters[n_blocks - 1],)
    return int(np.prod(preflatten)), preflatten


def encoder_spec(model_size, n_filters=(32, 64), n_blocks=2, activation="lrelu", batch_norm=True, kern_size=3,
                 kern_size_upconv=2, hidden_units=(128, 32, 2), pool_size=2, POOL_FLAG=True):
    spec, c = [], 1
    for i in range(n_blocks):
        c = _conv(spec, c, n_filters[i], kern_size, batch_norm)
        c = _conv(spec, c, 2 * n_filters[i], kern_size, batch_norm)
    flat, _ = encoder_shapes(model_size, n_filters, n_blocks, pool_size, POOL_FLAG)
    nin = flat
    for n_hidden in list(hidden_units)[:-1]:
        nin = _dense(spec, nin, n_hidden, batch_norm)
    _dense(spec, nin, hidden_units[-1], True)        # latent layer: BN always (autoencoders.py:339-341)
    return spec


def decoder_spec(model_size, n_filters=(32, 64), n_blocks=2, activation="lrelu", batch_norm=True, kern_size=3,
                 kern_size_upconv=2, hidden_units=(128, 32, 2), pool_size=2, POOL_FLAG=True):
    pools = [pool_size] * n_blocks if isinstance(pool_size, (int, np.integer)) else list(pool_size)
    if any(int(p) < 2 for p in pools[:n_blocks]):
        # Conv3DTranspose(kern_size_upconv = 2, strides = 1) overlaps its taps (autoencoders.py:213); not implemented
        raise NotImplementedError("decoder blocks need pool_size >= 2")
    flat, preflatten = encoder_shapes(model_size, n_filters, n_blocks, pool_size, POOL_FLAG)
    spec = []
    nin = hidden_units[-1]
    for n_hidden in list(hidden_units)[::-1][1:]:
        nin = _dense(spec, nin, n_hidden, batch_norm)
    _dense(spec, nin, flat, batch_norm)
    c = preflatten[-1]
    if POOL_FLAG:
        c = _upconv(spec, c, kern_size_upconv)
    for i in range(n_blocks - 1, -1, -1):
        c = _upconv(spec, c, kern_size_upconv)
        c = _conv(spec, c, n_filters[i], kern_size, batch_norm)
        c = _conv(spec, c, n_filters[i], kern_size, batch_norm)
    spec.append(("conv_k", (1, 1, 1, c, 1)))
    spec.append(("conv_b", (1,)))
    return spec


def keras_default_init(spec, seed=None):
    """Keras defaults: glorot_uniform kernels, zero biases, BN gamma=1 beta=0 mean=0 var=1."""
    rng = np.random.default_rng(seed)
    out = []
    for kind, shape in spec:
        if kind == "conv_k":
            rf = int(np.prod(shape[:3]))
            lim = np.sqrt(6.0 / (rf * shape[3] + rf * shape[4]))
            out.append(rng.uniform(-lim, lim, shape).astype(np.float32))
        elif kind == "convT_k":
            rf = int(np.prod(shape[:3]))
            lim = np.sqrt(6.0 / (rf * shape[4] + rf * shape[3]))
            out.append(rng.uniform(-lim, lim, shape).astype(np.float32))
        elif kind == "dense_k":
            lim = np.sqrt(6.0 / (shape[0] + shape[1]))
            out.append(rng.uniform(-lim, lim, shape).astype(np.float32))
        elif kind in ("bn_gamma", "bn_var"):
            out.append(np.ones(shape, np.float32))
        else:
            out.append(np.zeros(shape, np.float32))
    return out
