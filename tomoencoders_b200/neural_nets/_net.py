"""Thin Python handle over ``te_net`` (include/tomoenc.h): one Keras model of the reference
(U-net / encoder / decoder) planned, weighted and executed by libtomoenc.so."""
import ctypes

import numpy as np
import torch

from .. import _device, _lib

KINDS = {"unet": 0, "encoder": 1, "decoder": 2}


def _list8(values, n, name):
    values = list(values)
    if len(values) != n:
        raise ValueError("%s: list length must be equal to number of blocks" % name)
    if n > 8:
        raise ValueError("%s: at most 8 entries supported" % name)
    return values + [0] * (8 - n)


class Net:
    """kind in {"unet","encoder","decoder"}; model_params as in the reference builders
    (neural_nets/Unet3D.py:199-202, neural_nets/autoencoders.py:246-249)."""

    def __init__(self, kind, patch, max_batch, mode="bf16", n_filters=(16, 32, 64), n_blocks=None,
                 activation="lrelu", batch_norm=True, kern_size=3, kern_size_upconv=2, isconcat=None,
                 pool_size=2, hidden_units=(), POOL_FLAG=True):
        _device.require_cuda()
        if activation != "lrelu":
            raise NotImplementedError("only activation='lrelu' is implemented (the value of every shipped model)")
        n_filters = list(n_filters)
        n_blocks = len(n_filters) if n_blocks is None else int(n_blocks)
        if isinstance(pool_size, (int, np.integer)):
            pool_size = [int(pool_size)] * n_blocks
        elif len(pool_size) != n_blocks:
            raise ValueError("list length must be equal to number of blocks")
        if isconcat is None:
            isconcat = [False] * n_blocks
        d = _lib.NetDesc()
        d.kind = KINDS[kind]
        d.n_blocks = n_blocks
        d.n_filters[:] = _list8(n_filters[:n_blocks], n_blocks, "n_filters")
        d.pool_size[:] = _list8(pool_size, n_blocks, "pool_size")
        d.isconcat[:] = _list8([1 if c else 0 for c in isconcat], n_blocks, "isconcat")
        d.batch_norm = 1 if batch_norm else 0
        d.kern_size = int(kern_size)
        d.kern_size_upconv = int(kern_size_upconv)
        hidden_units = list(hidden_units)
        d.n_hidden = len(hidden_units)
        d.hidden_units[:] = hidden_units + [0] * (8 - len(hidden_units))
        d.pool_flag = 1 if POOL_FLAG else 0
        d.patch[:] = [int(p) for p in patch]
        d.max_batch = int(max_batch)
        d.lrelu_alpha = 0.2          # Unet3D.py:38, autoencoders.py:55
        d.bn_eps = 1e-5              # Unet3D.py:78, autoencoders.py:86
        d.mode = _lib.MODES[mode] if isinstance(mode, str) else int(mode)
        self.kind, self.mode, self.patch, self.max_batch = kind, d.mode, tuple(int(p) for p in patch), int(max_batch)
        self.desc = d
        self._L = _lib.lib()
        n = self._L.te_net_weight_count(ctypes.byref(d))
        if n < 0:
            _lib.check(int(n), "te_net_weight_count")
        self.weight_count = int(n)
        h = ctypes.c_void_p()
        _lib.check(self._L.te_net_create(ctypes.byref(d), ctypes.byref(h)), "te_net_create")
        self._h = h
        self.latent = hidden_units[-1] if hidden_units else 0
        self.act_dtype = {0: torch.bfloat16, 1: torch.float16, 2: torch.float32}[d.mode]
        self.act_te = {0: _lib.TE_BF16, 1: _lib.TE_F16, 2: _lib.TE_F32}[d.mode]
        self.profiling = False

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            self._L.te_net_destroy(h)
            self._h = None

    # ------------------------------------------------------------------ weights
    def set_weights(self, weights):
        """weights: list of arrays in Keras get_weights() order, or one flat float32 array."""
        if isinstance(weights, (list, tuple)):
            flat = np.concatenate([np.asarray(w, dtype=np.float32).reshape(-1) for w in weights])
        else:
            flat = np.ascontiguousarray(weights, dtype=np.float32).reshape(-1)
        if flat.size != self.weight_count:
            raise ValueError("expected %d weight values, got %d" % (self.weight_count, flat.size))
        flat = np.ascontiguousarray(flat)
        _lib.check(self._L.te_net_set_weights(self._h, flat.ctypes.data, flat.size), "te_net_set_weights")

    # ------------------------------------------------------------------ info
    @property
    def flops_per_patch(self):
        return float(self._L.te_net_flops_per_patch(self._h))

    @property
    def workspace_bytes(self):
        return int(self._L.te_net_workspace_bytes(self._h))

    def layers(self):
        out = []
        for i in range(self._L.te_net_num_layers(self._h)):
            name, fl, tc = ctypes.c_char_p(), ctypes.c_double(), ctypes.c_int()
            _lib.check(self._L.te_net_layer_info(self._h, i, ctypes.byref(name), ctypes.byref(fl), ctypes.byref(tc)))
            out.append((name.value.decode(), fl.value, bool(tc.value)))
        return out

    def overflowed(self, reset=True):
        """True when a forward call since the last reset produced an activation outside the finite fp16 range
        (te_net_overflow; always False for bf16 / fp32 handles).  Synchronises the current stream."""
        if self.act_dtype != torch.float16:
            return False
        flag = ctypes.c_int(0)
        _lib.check(self._L.te_net_overflow(self._h, 1 if reset else 0, ctypes.byref(flag), _device.stream_ptr()),
                   "te_net_overflow")
        return bool(flag.value)

    def set_profiling(self, enable):
        _lib.check(self._L.te_net_set_profiling(self._h, 1 if enable else 0), "te_net_set_profiling")
        self.profiling = bool(enable)

    def layer_times(self):
        """[(name, kernel, flops_per_patch, total_ms, launches)] accumulated since set_profiling(True)."""
        out = []
        buf = ctypes.create_string_buffer(128)
        for i, (name, fl, tc) in enumerate(self.layers()):
            ms, cnt = ctypes.c_double(), ctypes.c_int64()
            _lib.check(self._L.te_net_layer_time(self._h, i, ctypes.byref(ms), ctypes.byref(cnt)), "te_net_layer_time")
            _lib.check(self._L.te_net_layer_kernel(self._h, i, buf, 128), "te_net_layer_kernel")
            out.append((name, buf.value.decode(), fl, ms.value, cnt.value))
        return out

    def layer_output(self, index, shape):
        """fp32 NDHWC copy of a layer's activation tensor from the last forward call (test hook)."""
        out = torch.empty(shape, dtype=torch.float32, device=_device.device())
        n = self._L.te_net_layer_output(self._h, index, out.data_ptr(), out.numel(), _device.stream_ptr())
        if n < 0:
            _lib.check(int(n), "te_net_layer_output")
        if n != out.numel():
            raise ValueError("layer %d holds %d values, shape %s given" % (index, n, shape))
        return out

    # ------------------------------------------------------------------ forward (device tensors)
    def _check_x(self, x):
        if x.dtype != self.act_dtype or not x.is_cuda or not x.is_contiguous():
            raise TypeError("x must be a contiguous CUDA tensor of dtype %s" % self.act_dtype)
        if tuple(x.shape[1:4]) != self.patch or x.shape[0] > self.max_batch:
            raise ValueError("x must be (n <= %d,) + %s" % (self.max_batch, self.patch))

    def unet_forward(self, x, out_dtype=torch.uint8, want_logits=False):
        self._check_x(x)
        n = x.shape[0]
        out = torch.empty((n,) + self.patch, dtype=out_dtype, device=x.device)
        logits = torch.empty((n,) + self.patch, dtype=torch.float32, device=x.device) if want_logits else None
        code = _lib.TE_U8 if out_dtype == torch.uint8 else _lib.TE_F32
        _lib.check(self._L.te_unet_forward(self._h, x.data_ptr(), n, out.data_ptr(), code,
                                           logits.data_ptr() if want_logits else None, _device.stream_ptr()),
                   "te_unet_forward")
        return (out, logits) if want_logits else out

    def unet_forward_into(self, x, out):
        """Labels / probabilities written into a caller-owned (n,)+patch uint8 / float32 tensor."""
        self._check_x(x)
        code = _lib.TE_U8 if out.dtype == torch.uint8 else _lib.TE_F32
        _lib.check(self._L.te_unet_forward(self._h, x.data_ptr(), x.shape[0], out.data_ptr(), code, None,
                                           _device.stream_ptr()), "te_unet_forward")

    def encoder_forward(self, x, out=None):
        self._check_x(x)
        n = x.shape[0]
        z = torch.empty((n, self.latent), dtype=torch.float32, device=x.device) if out is None else out
        _lib.check(self._L.te_encoder_forward(self._h, x.data_ptr(), n, z.data_ptr(), _device.stream_ptr()),
                   "te_encoder_forward")
        return z

    def decoder_forward(self, z):
        if z.dtype != torch.float32 or not z.is_cuda or not z.is_contiguous():
            raise TypeError("z must be a contiguous float32 CUDA tensor")
        n = z.shape[0]
        out = torch.empty((n,) + self.patch, dtype=torch.float32, device=z.device)
        _lib.check(self._L.te_decoder_forward(self._h, z.data_ptr(), n, out.data_ptr(), _device.stream_ptr()),
                   "te_decoder_forward")
        return out
