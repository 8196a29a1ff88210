"""Voxel pre-/post-processing around the hot path: drop-in for the array functions of
/root/reference/tomo_encoders/misc/voxel_processing.py.

Everything runs on the GPU through the C ABI (csrc/gather_scatter.cu, csrc/voxel_ops.cu); there is
no CPU fallback.  numpy in -> numpy out (host<->device copies inside the call), torch CUDA tensor in
-> torch CUDA tensor out, like the reference's numpy / cupy duality (``cp.get_array_module``).
Inside predict_patches / predict_embeddings the clip + rescale is fused into the gather kernel
instead (te_gather_norm).
"""
import math
import time

import numpy as np
import torch

from .. import _device, _lib

_WHOLE_VOLUME_BYTES = 32 << 30      # host volumes up to this size are normalised in one upload


class TimerGPU:
    """voxel_processing.py:29-51: a CUDA-event pair on the current stream; toc() returns the elapsed time in `unit`
    ("ms" or "secs") and prints it when a message is given."""

    def __init__(self, unit):
        self.unit = unit

    def tic(self):
        _device.require_cuda()
        self.start = torch.cuda.Event(enable_timing=True)
        self.end = torch.cuda.Event(enable_timing=True)
        self.start.record()

    def toc(self, msg="execution"):
        self.end.record()
        self.end.synchronize()
        t_elapsed = self.start.elapsed_time(self.end)
        if self.unit == "secs":
            t_elapsed /= 1000.0
        if msg != "execution":
            print(f"\tTIME: {msg} {t_elapsed:.2f} {self.unit}")
        return t_elapsed


class TimerCPU:
    """voxel_processing.py:53-70: wall-clock timer; unit "ms" or "secs"."""

    def __init__(self, unit):
        self.unit = unit

    def tic(self):
        self.start = time.time()

    def toc(self, msg="execution"):
        t_elapsed = time.time() - self.start
        if self.unit == "ms":
            t_elapsed *= 1000.0
        if msg != "execution":
            print(f"\tTIME: {msg} {t_elapsed:.2f} {self.unit}")
        return t_elapsed


def _msg_exec_time(func, t_exec):
    """voxel_processing.py:72-77."""
    print("TIME: %s: %.2f seconds" % (func if type(func) is str else func.__name__, t_exec))


def _rescale_data(data, min_val, max_val):
    """voxel_processing.py:81-89: (data - min) / (max - min + 1e-12).

    Host arrays keep the reference's numpy arithmetic (dtype promotion included).  Device tensors of
    a supported element type run te_rescale: float32 arithmetic, float32 result (float16 stays
    float16), bit-exact against numpy for float32 input."""
    eps = 1e-12
    if _device.is_device_array(data) and data.dtype in (torch.uint8, torch.uint16, torch.float16, torch.float32):
        t = data.contiguous()
        out_dtype = torch.float16 if t.dtype == torch.float16 else torch.float32
        out = torch.empty(t.shape, dtype=out_dtype, device=t.device)
        lo = np.float32(min_val)
        denom = np.float32(np.float32(max_val) - lo + np.float32(eps))
        _lib.check(_lib.lib().te_rescale(t.data_ptr(), _device.te_dtype(t), out.data_ptr(), _device.te_dtype(out),
                                         t.numel(), float(lo), float(denom), _device.stream_ptr()), "te_rescale")
        return out
    return (data - min_val) / (max_val - min_val + eps)


def _find_min_max(vol, sampling_factor, TIMEIT=False):
    """voxel_processing.py:91-104: (min, max) of vol[::s, ::s, ::s]; returns Python floats."""
    t0 = time.time()
    s = int(sampling_factor)
    h = None if _device.is_device_array(vol) else _device.as_host_tensor(vol)
    if h is not None and h.dim() == 3 and s > 1 and h.shape[0] >= s and h.is_contiguous():
        # host volume: vol[::s, ::s, ::s] needs every s-th row of every s-th plane only -- ONE strided 3-D DMA moves 1 / s^2
        # of the bytes (te_upload_sampled), the x stride is applied by the device-side min / max
        dev = _device.device()
        planes = torch.empty((-(-h.shape[0] // s), -(-h.shape[1] // s), h.shape[2]), dtype=h.dtype, device=dev)
        _lib.check(_lib.lib().te_upload_sampled(h.data_ptr(), h.element_size(), _lib.i64x3(h.shape), s, planes.data_ptr(),
                                                _device.stream_ptr()), "te_upload_sampled")
        lo, hi = _minmax_view(planes[:, :, ::s])
        if TIMEIT:
            _msg_exec_time("find voxel min max", time.time() - t0)
        return float(lo), float(hi)
    t = _device.to_supported(_device.to_device(vol))
    if t.dim() != 3:
        raise ValueError("expected a 3-D volume")
    out = torch.empty(2, dtype=torch.float32, device=t.device)
    scratch = torch.empty(2048, dtype=torch.float32, device=t.device)
    L = _lib.lib()
    _lib.check(L.te_minmax_strided(t.data_ptr(), _device.te_dtype(t), _lib.i64x3(t.shape), int(sampling_factor),
                                   out.data_ptr(), scratch.data_ptr(), _device.stream_ptr()), "te_minmax_strided")
    lo, hi = out.tolist()
    if TIMEIT:
        _msg_exec_time("find voxel min max", time.time() - t0)
    return lo, hi


# ---------------------------------------------------------------------------------------------- views
def _as_device(a):
    """numpy / torch array of a supported element type (bool counts as uint8) -> torch CUDA tensor
    (strided views of device tensors are kept as views: the kernels take element strides)."""
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(_device.device())
    else:
        a = np.asarray(a)
        if a.dtype == np.bool_:
            a = a.view(np.uint8)
        t = _device.to_device(a)
    if t.dtype == torch.bool:
        t = t.view(torch.uint8)
    return t


def _view4(t):
    """(shape[4], strides[4]) ctypes arrays of a <= 4-D tensor, outer dimensions padded with 1."""
    if t.dim() > 4:
        raise ValueError("at most 4 dimensions are supported")
    pad = 4 - t.dim()
    shape = [1] * pad + [int(s) for s in t.shape]
    strides = [0] * pad + [int(s) for s in t.stride()]
    I4 = _lib._I64 * 4
    return I4(*shape), I4(*strides)


def _minmax_view(t):
    """Exact (min, max) of a strided device view as float32 values."""
    out = torch.empty(2, dtype=torch.float32, device=t.device)
    scratch = torch.empty(2048, dtype=torch.float32, device=t.device)
    shape, strides = _view4(t)
    _lib.check(_lib.lib().te_minmax_nd(t.data_ptr(), _device.te_dtype(t), shape, strides, out.data_ptr(),
                                       scratch.data_ptr(), _device.stream_ptr()), "te_minmax_nd")
    lo, hi = out.tolist()
    return np.float32(lo), np.float32(hi)


_NP_OF_TORCH = {torch.uint8: np.uint8, torch.uint16: np.uint16, torch.float16: np.float16, torch.float32: np.float32}


def _histogram_view(t, bins):
    """np.histogram(view, bins=bins) on the device: (counts int64, edges) with numpy's own edges."""
    if t.numel() == 0:
        raise ValueError("histogram of an empty array")
    lo, hi = _minmax_view(t)
    np_dt = _NP_OF_TORCH[t.dtype]
    # numpy derives the edges from (a.min(), a.max(), a.dtype) only: ask numpy itself
    edges = np.histogram_bin_edges(np.array([lo, hi]).astype(np_dt), bins=bins)
    d_edges = torch.from_numpy(np.ascontiguousarray(edges, dtype=np.float64)).to(t.device)
    counts = torch.empty(bins, dtype=torch.int64, device=t.device)
    shape, strides = _view4(t)
    _lib.check(_lib.lib().te_histogram_nd(t.data_ptr(), _device.te_dtype(t), shape, strides, d_edges.data_ptr(),
                                          int(bins), counts.data_ptr(), _device.stream_ptr()), "te_histogram_nd")
    return counts.cpu().numpy(), edges


def modified_autocontrast(vol, s=0.01, normalize_sampling_factor=2):
    """voxel_processing.py:160-204: (alow, ahigh) clamp values saturating the lowest / highest
    quantile ``s`` (or ``(slow, shigh)``) of vol[::f, ::f, ::f] (4-D: vol[:, ::f, ::f, ::f]; 1-D: all
    values), read off a 500-bin histogram.  The histogram is accumulated on the GPU bit-exactly
    against np.histogram; the 500-entry CDF search stays on the host.  Returns numpy scalars in the
    bin dtype numpy would produce."""
    t = _as_device(vol)
    f = int(normalize_sampling_factor)
    if t.dim() == 4:
        view = t[:, ::f, ::f, ::f]
    elif t.dim() == 3:
        view = t[::f, ::f, ::f]
    elif t.dim() == 1:
        view = t
    else:
        raise ValueError("vol must be 1-, 3- or 4-dimensional")
    if type(s) == tuple and len(s) == 2:
        slow, shigh = s
    else:
        slow = shigh = s
    h, bins = _histogram_view(view, 500)
    c = np.cumsum(h)
    c_norm = c / np.max(c)
    ibin_low = np.argmin(np.abs(c_norm - slow))
    ibin_high = np.argmin(np.abs(c_norm - 1 + shigh))
    return bins[ibin_low], bins[ibin_high]


def _rescale_inplace(t, lo, hi):
    denom = np.float32(np.float32(hi) - np.float32(lo) + np.float32(1e-12))
    _lib.check(_lib.lib().te_rescale(t.data_ptr(), _device.te_dtype(t), t.data_ptr(), _device.te_dtype(t), t.numel(),
                                     float(lo), float(denom), _device.stream_ptr()), "te_rescale")


def normalize_volume_gpu(vol, chunk_size=64, normalize_sampling_factor=1, TIMEIT=False):
    """voxel_processing.py:207-253: rescales ``vol`` IN PLACE to [0, 1] with the min / max of
    vol[::f, ::f, ::f] and returns it.  float32 (the only dtype the reference's cupy staging buffer
    accepts) is bit-exact against numpy float32 arithmetic; float16 is computed in float32 and
    rounded.  Device tensors are rescaled where they live; host arrays go through the GPU whole
    (<= 32 GiB) or in ``chunk_size`` z-slabs (two passes: min / max, then rescale)."""
    t0 = time.time()
    f = int(normalize_sampling_factor)
    dt = vol.dtype
    if dt not in (np.float32, np.float16, torch.float32, torch.float16):
        raise TypeError("normalize_volume_gpu needs a float32 (or float16) volume, got %s" % (dt,))
    if _device.is_device_array(vol):
        if not vol.is_contiguous():
            raise ValueError("device volumes must be contiguous")
        lo, hi = _minmax_view(vol[::f, ::f, ::f])
        _rescale_inplace(vol, lo, hi)
    else:
        host = vol if isinstance(vol, torch.Tensor) else torch.from_numpy(vol)
        nz = host.shape[0]
        if host.numel() * host.element_size() <= _WHOLE_VOLUME_BYTES:
            d = host.to(_device.device())
            lo, hi = _minmax_view(d[::f, ::f, ::f])
            _rescale_inplace(d, lo, hi)
            host.copy_(d)
        else:
            lo, hi = np.float32(np.inf), np.float32(-np.inf)
            for z0 in range(0, nz, chunk_size):
                d = host[z0:z0 + chunk_size].to(_device.device())
                sub = d[(-z0) % f::f, ::f, ::f]
                if sub.numel():
                    a, b = _minmax_view(sub)
                    lo, hi = min(lo, a), max(hi, b)
            for z0 in range(0, nz, chunk_size):
                d = host[z0:z0 + chunk_size].to(_device.device())
                _rescale_inplace(d, lo, hi)
                host[z0:z0 + chunk_size].copy_(d)
    if TIMEIT:
        print("time for normalizing volume: %.2f secs" % float(time.time() - t0))
    return vol


def edge_map(Y):
    """voxel_processing.py:106-122: boolean volume, True where a voxel differs from any of its six
    neighbours (te_edge_map: one pass, 128-bit loads, SIMD-in-register comparisons)."""
    t = _as_device(Y).contiguous()
    if t.dim() != 3:
        raise ValueError("expected a 3-D volume")
    out = torch.empty(t.shape, dtype=torch.bool, device=t.device)
    _lib.check(_lib.lib().te_edge_map(t.data_ptr(), t.element_size(), _lib.i64x3(t.shape), out.data_ptr(),
                                      _device.stream_ptr()), "te_edge_map")
    return _device.to_host_like(out, Y)


def _cyl_radius(vol_shape, mask_fac):
    assert vol_shape[1] == vol_shape[2], "must be a tomographic volume where shape y = shape x"
    n = int(vol_shape[1])
    if n % 2:
        # the reference builds its circle from arange(-(n//2), ceil(n//2)): n - 1 points for odd n, and
        # the boolean index then fails on the shape mismatch
        raise IndexError("boolean index did not match indexed array: y / x extent %d must be even" % n)
    return n, int(mask_fac * n / 2)


def cylindrical_mask(out_vol, mask_fac, mask_val=0):
    """voxel_processing.py:124-140: sets, in place, every voxel outside the inscribed cylinder of
    relative radius ``mask_fac`` to ``mask_val``.  Returns None."""
    n, rad = _cyl_radius(out_vol.shape, mask_fac)
    on_device = _device.is_device_array(out_vol)
    t = out_vol if on_device else _as_device(out_vol)
    if not t.is_contiguous():
        raise ValueError("device volumes must be contiguous")
    np_dt = np.dtype(str(t.dtype).replace("torch.", "")) if t.dtype != torch.bool else np.dtype(np.bool_)
    cell = np.empty(1, dtype=np_dt)
    cell[0] = mask_val.item() if isinstance(mask_val, torch.Tensor) else mask_val      # numpy's assignment cast
    bits = int(cell.view({1: np.uint8, 2: np.uint16, 4: np.uint32}[np_dt.itemsize])[0])
    tv = t.view(torch.uint8) if t.dtype == torch.bool else t
    _lib.check(_lib.lib().te_cylindrical_mask(tv.data_ptr(), tv.element_size(), _lib.i64x3(tv.shape), rad, bits,
                                              _device.stream_ptr()), "te_cylindrical_mask")
    if not on_device:
        if isinstance(out_vol, torch.Tensor):
            out_vol.copy_(t)
        else:
            out_vol[...] = t.cpu().numpy().view(out_vol.dtype)
    return


_CYL_ROWS = {}


def _cyl_rows(n, rad, device):
    """Per-row inside interval [x_lo, x_hi) and exclusive prefix sum of its lengths."""
    key = (n, rad, device.index)
    hit = _CYL_ROWS.get(key)
    if hit is None:
        pts = np.arange(-(n // 2), math.ceil(n // 2), dtype=np.int64)
        r2 = pts * pts
        inside = (r2[:, None] + r2[None, :]) < (rad * rad if rad > 0 else 0)
        cnt = inside.sum(axis=1).astype(np.int64)
        x_lo = np.where(cnt > 0, inside.argmax(axis=1), 0).astype(np.int32)
        x_hi = (x_lo + cnt).astype(np.int32)
        off = np.concatenate([[0], np.cumsum(cnt)[:-1]]).astype(np.int64)
        hit = (torch.from_numpy(x_lo).to(device), torch.from_numpy(x_hi).to(device), torch.from_numpy(off).to(device),
               int(cnt.sum()))
        if len(_CYL_ROWS) > 8:
            _CYL_ROWS.clear()
        _CYL_ROWS[key] = hit
    return hit


def get_values_cyl_mask(vol, mask_fac):
    """voxel_processing.py:142-157: the voxels inside the cylinder as a 1-D array in C order
    (``vol[cyl > 0]``).  Sub-sampled device views (V[::2, ::2, ::2]) are read in place."""
    n, rad = _cyl_radius(vol.shape, mask_fac)
    t = _as_device(vol)
    if t.dim() != 3:
        raise ValueError("expected a 3-D volume")
    x_lo, x_hi, off, per_slice = _cyl_rows(n, rad, t.device)
    out = torch.empty(int(t.shape[0]) * per_slice, dtype=t.dtype, device=t.device)
    _lib.check(_lib.lib().te_cyl_values(t.data_ptr(), t.element_size(), _lib.i64x3(t.shape), _lib.i64x3(t.stride()),
                                        x_lo.data_ptr(), x_hi.data_ptr(), off.data_ptr(), per_slice, out.data_ptr(),
                                        _device.stream_ptr()), "te_cyl_values")
    if _device.is_device_array(vol):
        return out.view(torch.bool) if vol.dtype == torch.bool else out
    res = out.cpu().numpy()
    src_dt = vol.dtype if not isinstance(vol, torch.Tensor) else None
    return res.view(np.bool_) if src_dt == np.bool_ else res
