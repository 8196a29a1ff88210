"""torch-CPU restatement of the reference's Keras graphs.  TEST INFRASTRUCTURE.

PARITY UNPINNED: TensorFlow cannot be installed in this image and the reference keeps no golden
vectors for network outputs, so these functions restate the Keras layer
semantics documented in SURVEY.md appendix A.3 and follow the reference builders line by line:

  U-net      /root/reference/tomo_encoders/neural_nets/Unet3D.py:44-291
  encoder    /root/reference/tomo_encoders/neural_nets/autoencoders.py:60-91,146-189,246-347
  decoder    /root/reference/tomo_encoders/neural_nets/autoencoders.py:192-243,350-447
  wrappers   neural_nets/keras_processor.py:78-146, surface_segmenter.py:246-294,
             autoencoders.py:729-777

Tensors are NCDHW inside torch; inputs/outputs of the public functions are NDHWC numpy arrays like
Keras'.  Weights come as a flat list in Keras ``get_weights()`` order (oracle/weights.py).
"""
import numpy as np
import torch
import torch.nn.functional as F

LRELU_ALPHA = 0.2  # Unet3D.py:38, autoencoders.py:55
BN_EPS = 1e-5      # Unet3D.py:78, autoencoders.py:86,138


class _W:
    """Pops weights in creation order; in spec mode records (kind, shape) instead."""

    def __init__(self, weights=None, dtype=torch.float32):
        self.weights = weights
        self.i = 0
        self.spec = []
        self.dtype = dtype

    @property
    def spec_mode(self):
        return self.weights is None

    def take(self, kind, shape):
        self.spec.append((kind, tuple(int(s) for s in shape)))
        if self.spec_mode:
            return None
        w = self.weights[self.i]
        self.i += 1
        assert tuple(w.shape) == tuple(shape), (kind, w.shape, shape)
        return torch.from_numpy(np.ascontiguousarray(w)).to(self.dtype)


class _T:
    """Shape-only stand-in for a tensor in spec mode."""

    def __init__(self, c, size):
        self.c, self.size = c, tuple(size)


def _channels(t):
    return t.c if isinstance(t, _T) else t.shape[1]


def _lrelu(t):
    return t if isinstance(t, _T) else F.leaky_relu(t, LRELU_ALPHA)


def _bn(t, W, c):
    g = W.take("bn_gamma", (c,))
    b = W.take("bn_beta", (c,))
    m = W.take("bn_mean", (c,))
    v = W.take("bn_var", (c,))
    if isinstance(t, _T):
        return t
    shape = (1, c) + (1,) * (t.dim() - 2)
    return g.view(shape) * (t - m.view(shape)) / torch.sqrt(v.view(shape) + BN_EPS) + b.view(shape)


def custom_conv3d(t, W, n_filters, kern_size, activation="lrelu", batch_norm=False):
    """Unet3D.py:44-81 / autoencoders.py:106-141: Conv3D(same) -> [BN] -> activation."""
    cin = _channels(t)
    k = W.take("conv_k", (kern_size,) * 3 + (cin, n_filters))
    b = W.take("conv_b", (n_filters,))
    if isinstance(t, _T):
        out = _T(n_filters, t.size)
    else:
        out = F.conv3d(t, k.permute(4, 3, 0, 1, 2).contiguous(), b, padding=kern_size // 2)
    if batch_norm:
        out = _bn(out, W, n_filters)
    return _lrelu(out) if activation == "lrelu" else out


def _maxpool(t, p):
    """MaxPool3D(pool, padding='same') for sizes divisible by the pool."""
    if isinstance(t, _T):
        assert all(s % p == 0 for s in t.size), "pool must divide the spatial size"
        return _T(t.c, [s // p for s in t.size])
    assert all(s % p == 0 for s in t.shape[2:]), "pool must divide the spatial size"
    return F.max_pool3d(t, p, p)


def _upconv(t, W, k, stride):
    """Conv3DTranspose(C, k, strides=stride, padding='same') + LeakyReLU (Unet3D.py:175-177).
    Output length = in*stride; for k < stride the gap voxels get bias only."""
    c = _channels(t)
    kern = W.take("convT_k", (k,) * 3 + (c, c))
    b = W.take("convT_b", (c,))
    if isinstance(t, _T):
        return _T(c, [s * stride for s in t.size])
    assert k <= stride, "restated for non-overlapping transposed convs only (k <= stride)"
    out = F.conv_transpose3d(t, kern.permute(4, 3, 0, 1, 2).contiguous(), b, stride=stride,
                             output_padding=stride - k)
    return _lrelu(out)


def _concat(a, b):
    if isinstance(a, _T):
        return _T(a.c + b.c, a.size)
    return torch.cat([a, b], dim=1)


# ------------------------------------------------------------------------------------ U-net
def unet_graph(t, W, n_filters=(16, 32, 64), n_blocks=3, activation="lrelu", batch_norm=True,
               kern_size=3, kern_size_upconv=2, isconcat=None, pool_size=2, return_logits=False):
    """build_Unet_3D, Unet3D.py:199-291."""
    if isconcat is None:
        isconcat = [False] * n_blocks
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    elif len(pool_size) != n_blocks:
        raise ValueError("list length must be equal to number of blocks")
    concats = []
    code = t
    for ii in range(n_blocks):                                    # analysis_block :84-132
        code = custom_conv3d(code, W, n_filters[ii], kern_size, activation, batch_norm)
        code = custom_conv3d(code, W, 2 * n_filters[ii], kern_size, activation, batch_norm)
        concats.append(code)
        if pool_size[ii] > 1:
            code = _maxpool(code, pool_size[ii])
    nf = _channels(code)                                          # :267-271
    code = custom_conv3d(code, W, nf, kern_size, activation, batch_norm)
    dec = custom_conv3d(code, W, 2 * nf, kern_size, activation, batch_norm)
    for ii in range(n_blocks - 1, -1, -1):                        # synthesis_block :134-194
        if pool_size[ii] > 1:
            dec = _upconv(dec, W, kern_size_upconv, pool_size[ii])
            if isconcat[ii]:
                dec = _concat(dec, concats[ii])                   # up first, skip second (:180)
        dec = custom_conv3d(dec, W, 2 * n_filters[ii], kern_size, activation, batch_norm)
        dec = custom_conv3d(dec, W, 2 * n_filters[ii], kern_size, activation, batch_norm)
    cin = _channels(dec)                                          # :287, 1x1x1 conv + sigmoid
    k = W.take("conv_k", (1, 1, 1, cin, 1))
    b = W.take("conv_b", (1,))
    if isinstance(dec, _T):
        return dec
    logits = F.conv3d(dec, k.permute(4, 3, 0, 1, 2).contiguous(), b)
    return logits if return_logits else torch.sigmoid(logits)


def unet_spec(**model_params):
    W = _W(None)
    unet_graph(_T(1, (64, 64, 64)), W, **model_params)
    return W.spec


def unet_forward(x, weights, dtype=torch.float32, return_logits=False, **model_params):
    """x: (N,D,H,W,1) numpy -> (N,D,H,W,1) numpy (sigmoid probabilities, or logits)."""
    W = _W(weights, dtype)
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype).permute(0, 4, 1, 2, 3)
    with torch.no_grad():
        y = unet_graph(t, W, return_logits=return_logits, **model_params)
    assert W.i == len(weights)
    return y.permute(0, 2, 3, 4, 1).contiguous().numpy()


# ------------------------------------------------------------------------------------ encoder / decoder
def _hidden_layer(t, W, n_hidden, activation="lrelu", batch_norm=False):
    """autoencoders.py:60-91: Dense -> [BN] -> activation."""
    cin = t.c if isinstance(t, _T) else t.shape[1]
    k = W.take("dense_k", (cin, n_hidden))
    b = W.take("dense_b", (n_hidden,))
    out = _T(n_hidden, ()) if isinstance(t, _T) else t @ k + b
    if batch_norm:
        out = _bn(out, W, n_hidden)
    return _lrelu(out) if activation == "lrelu" else out


def encoder_graph(t, W, n_filters=(32, 64), n_blocks=2, activation="lrelu", batch_norm=True,
                  kern_size=3, kern_size_upconv=2, hidden_units=(128, 32, 2), pool_size=2,
                  POOL_FLAG=True):
    """build_encoder_r, autoencoders.py:246-347.  Returns (z, flatten_shape, preflatten_shape)."""
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    elif len(pool_size) != n_blocks:
        raise ValueError("list length must be equal to number of blocks")
    code = t
    for ii in range(n_blocks):                                    # analysis_block_small :146-189
        code = custom_conv3d(code, W, n_filters[ii], kern_size, activation, batch_norm)
        code = custom_conv3d(code, W, 2 * n_filters[ii], kern_size, activation, batch_norm)
        code = _maxpool(code, pool_size[ii])
    if POOL_FLAG:
        code = _maxpool(code, 2)                                  # :317-319
    if isinstance(code, _T):
        preflatten = tuple(code.size) + (code.c,)
        flat = _T(int(np.prod(preflatten)), ())
    else:
        preflatten = tuple(code.shape[2:]) + (code.shape[1],)
        flat = code.permute(0, 2, 3, 4, 1).reshape(code.shape[0], -1)   # Keras Flatten of NDHWC
    flatten_shape = int(np.prod(preflatten))
    code = flat
    for ic, n_hidden in enumerate(hidden_units):                  # :321-337
        if ic == len(hidden_units) - 1:
            break
        code = _hidden_layer(code, W, n_hidden, activation, batch_norm)
    z = _hidden_layer(code, W, hidden_units[-1], activation, True)  # :339-341, BN always
    return z, flatten_shape, preflatten


def encoder_spec(model_size, **model_params):
    W = _W(None)
    _, flatten_shape, preflatten = encoder_graph(_T(1, model_size), W, **model_params)
    return W.spec, flatten_shape, preflatten


def encoder_forward(x, weights, dtype=torch.float32, **model_params):
    """x: (N,D,H,W,1) -> (N, latent) float32."""
    W = _W(weights, dtype)
    t = torch.from_numpy(np.ascontiguousarray(x)).to(dtype).permute(0, 4, 1, 2, 3)
    with torch.no_grad():
        z, _, _ = encoder_graph(t, W, **model_params)
    assert W.i == len(weights)
    return z.numpy()


def decoder_graph(z, W, flatten_shape, preflatten_shape, n_filters=(32, 64), n_blocks=2,
                  activation="lrelu", batch_norm=True, kern_size=3, kern_size_upconv=2,
                  hidden_units=(128, 32, 2), pool_size=2, POOL_FLAG=True, return_logits=False):
    """build_decoder_r, autoencoders.py:350-447 (synthesis_block_small :192-243)."""
    if isinstance(pool_size, int):
        pool_size = [pool_size] * n_blocks
    dec = z
    for ic, n_hidden in enumerate(list(hidden_units)[::-1]):      # :407-414
        if ic == 0:
            continue
        dec = _hidden_layer(dec, W, n_hidden, activation, batch_norm)
    dec = _hidden_layer(dec, W, flatten_shape, activation, batch_norm)  # :417
    if isinstance(dec, _T):
        dec = _T(preflatten_shape[-1], preflatten_shape[:-1])
    else:
        dec = dec.reshape((dec.shape[0],) + tuple(preflatten_shape)).permute(0, 4, 1, 2, 3)
    if POOL_FLAG:
        dec = _upconv(dec, W, kern_size_upconv, 2)                # :422-430
    for ii in range(n_blocks - 1, -1, -1):
        dec = _upconv(dec, W, kern_size_upconv, pool_size[ii])
        dec = custom_conv3d(dec, W, n_filters[ii], kern_size, activation, batch_norm)
        dec = custom_conv3d(dec, W, n_filters[ii], kern_size, activation, batch_norm)
    cin = _channels(dec)
    k = W.take("conv_k", (1, 1, 1, cin, 1))
    b = W.take("conv_b", (1,))
    if isinstance(dec, _T):
        return dec
    logits = F.conv3d(dec, k.permute(4, 3, 0, 1, 2).contiguous(), b)
    return logits if return_logits else torch.sigmoid(logits)


def decoder_spec(model_size, **model_params):
    _, flatten_shape, preflatten = encoder_spec(model_size, **model_params)
    W = _W(None)
    decoder_graph(_T(model_params.get("hidden_units", (128, 32, 2))[-1], ()), W, flatten_shape,
                  preflatten, **model_params)
    return W.spec


def decoder_forward(z, weights, model_size, dtype=torch.float32, **model_params):
    _, flatten_shape, preflatten = encoder_spec(model_size, **model_params)
    W = _W(weights, dtype)
    with torch.no_grad():
        y = decoder_graph(torch.from_numpy(np.ascontiguousarray(z)).to(dtype), W, flatten_shape,
                          preflatten, **model_params)
    assert W.i == len(weights)
    return y.permute(0, 2, 3, 4, 1).contiguous().numpy()


# ------------------------------------------------------------------------------------ wrappers
def _rescale_reference(x_in, min_max, clip):
    """surface_segmenter.py:272-275 / keras_processor.py:102-104 with numpy's own dtype promotion."""
    if min_max is None:
        return x_in
    min_val, max_val = min_max
    if clip:
        x_in = np.clip(x_in, min_val, max_val)
    return (x_in - float(min_val)) / (float(max_val) - float(min_val) + 1e-12)


def _chunked(x, chunk_size, min_max, clip, fn, out_arr):
    """Chunk loop with edge padding of the last chunk (keras_processor.py:87-131)."""
    nb = len(x)
    nchunks = int(np.ceil(nb / chunk_size))
    padding = nchunks * chunk_size - nb
    for k in range(nchunks):
        sb = slice(k * chunk_size, min((k + 1) * chunk_size, nb))
        x_in = _rescale_reference(x[sb, ...], min_max, clip)
        if padding != 0 and k == nchunks - 1:
            x_in = np.pad(x_in, ((0, padding), (0, 0), (0, 0), (0, 0), (0, 0)), mode="edge")
            x_out = fn(np.asarray(x_in, dtype=np.float32))[:-padding, ...]
        else:
            x_out = fn(np.asarray(x_in, dtype=np.float32))
        out_arr[sb, ...] = x_out
    return out_arr


def segmenter_predict_patches(x, chunk_size, weights, min_max=None, out_arr=None, **model_params):
    """SurfaceSegmenter.predict_patches, surface_segmenter.py:246-294: clip, rescale, predict,
    np.round -> uint8."""
    assert x.ndim == 5
    if out_arr is None:
        out_arr = np.empty_like(x)
    _chunked(x, chunk_size, min_max, True, lambda xi: unet_forward(xi, weights, **model_params), out_arr)
    return np.round(out_arr).astype(np.uint8)


def segmenter_predict_probs(x, chunk_size, weights, min_max=None, **model_params):
    """Same pipeline but returns float32 probabilities (used to measure the decision margin)."""
    out = np.empty(x.shape, dtype=np.float32)
    return _chunked(x, chunk_size, min_max, True, lambda xi: unet_forward(xi, weights, **model_params), out)


def standardize(x):
    """``standardize`` / ``stdize_vol`` of the ICIP porosity encoder (scratchpad/IEEE_ICIP_paper/porosity_encoders.py:89-98,
    the Lambda in front of build_CAE_3D when stdinput=True, :298-302): every volume of the batch is mapped to
    (vol - min) / (max - min + 1e-12) in float32.  x: (N, D, H, W, 1)."""
    x = np.asarray(x, dtype=np.float32)
    ax = tuple(range(1, x.ndim))
    mn = x.min(axis=ax, keepdims=True)
    mx = x.max(axis=ax, keepdims=True)
    return ((x - mn) / (mx - mn + np.float32(1e-12))).astype(np.float32)


def encoder_predict_embeddings(x, chunk_size, weights, min_max=None, stdinput=False, **model_params):
    """SelfSupervisedCAE.predict_embeddings, autoencoders.py:729-777 (no clip).  The reference's
    NameError on ``_rescale_data`` when min_max is given (quirk Q4) is not reproduced."""
    assert x.ndim == 5
    W = _W(None)
    z, _, _ = encoder_graph(_T(1, x.shape[1:4]), W, **model_params)
    out = np.zeros((len(x), z.c), dtype=np.float32)
    fwd = (lambda xi: encoder_forward(standardize(xi), weights, **model_params)) if stdinput else \
        (lambda xi: encoder_forward(xi, weights, **model_params))
    return _chunked(x, chunk_size, min_max, False, fwd, out)
