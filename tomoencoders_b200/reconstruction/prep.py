"""Projection pre-processing: drop-in for /root/reference/tomo_encoders/reconstruction/prep.py:21-77."""
import numpy as np
import torch

from .. import _device, _lib
from .cuda_kernels import _Timer, _f32_dev


def calc_padding(data_shape):
    """prep.py:21-31: 1/4 padding on both sides, padded width a multiple of 8."""
    ntheta, nz, n = data_shape
    n_pad = n * (1 + 0.25 * 2)
    n_pad = int(np.ceil(n_pad / 8.0) * 8.0)
    pad_left = int((n_pad - n) // 2)
    pad_right = n_pad - n - pad_left
    return pad_left, pad_right


def _filter_dev(d):
    rows, n = int(d.shape[0]) * int(d.shape[1]), int(d.shape[2])
    L = _lib.lib()
    nbytes = int(L.te_fbp_workspace_bytes(rows, n))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=d.device)
    _lib.check(L.te_fbp_filter(d.data_ptr(), rows, n, ws.data_ptr(), nbytes, _device.stream_ptr()), "te_fbp_filter")


def fbp_filter(data):
    """prep.py:33-60: ramp filter (|f| = rfftfreq) of every detector row, in place; returns the elapsed milliseconds.
    data: float32 (ntheta, nz, n), a contiguous CUDA tensor (filtered in place) or a numpy array (copied both ways)."""
    t = _Timer()
    if _device.is_device_array(data):
        if data.dtype != torch.float32 or not data.is_contiguous() or data.dim() != 3:
            raise TypeError("data must be a contiguous float32 (ntheta, nz, n) tensor")
        _filter_dev(data)
    else:
        d = _f32_dev(data)
        if d.dim() != 3:
            raise TypeError("data must be (ntheta, nz, n)")
        _filter_dev(d)
        data[...] = d.cpu().numpy()
    return t.stop()


def _preprocess_dev(d, dark, flat):
    dk, fl = _f32_dev(dark), _f32_dev(flat)
    ntheta, nz, n = (int(s) for s in d.shape)
    if tuple(dk.shape) != (nz, n) or tuple(fl.shape) != (nz, n):
        raise ValueError("dark / flat must be (nz, n) = %s" % ((nz, n),))
    _lib.check(_lib.lib().te_preprocess(d.data_ptr(), dk.data_ptr(), fl.data_ptr(), ntheta, nz, n, _device.stream_ptr()),
               "te_preprocess")


def preprocess(data, dark, flat):
    """prep.py:63-77 in place: dark / flat correction and -log (the reference's median_filter([1,1,1]) is the identity and
    its outlier replacement therefore never fires)."""
    if _device.is_device_array(data):
        if data.dtype != torch.float32 or not data.is_contiguous() or data.dim() != 3:
            raise TypeError("data must be a contiguous float32 (ntheta, nz, n) tensor")
        _preprocess_dev(data, dark, flat)
    else:
        d = _f32_dev(data)
        _preprocess_dev(d, dark, flat)
        data[...] = d.cpu().numpy()
    return
