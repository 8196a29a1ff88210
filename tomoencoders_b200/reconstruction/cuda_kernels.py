"""Back-projection kernels: drop-in for /root/reference/tomo_encoders/reconstruction/cuda_kernels.py:162-302
(rec_pts, rec_pts_xy, rec_mask, rec_all, rec_patch) over te_bp_* (csrc/recon_ops.cu).

The reference takes CuPy arrays; here torch CUDA tensors are the device handle and numpy arrays are uploaded /
downloaded inside the call.  `exact=True` selects the parity mode of the kernels (one IEEE rounding per reference
operation); the default is the FMA-contracted product mode.  There is no CPU fallback.
"""
import numpy as np
import torch

from .. import _device, _lib


def _f32_dev(a):
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(_device.device())
        return t.to(torch.float32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(_device.device())


class _Timer:
    def __init__(self):
        self.a, self.b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.a.record()

    def stop(self):
        self.b.record()
        self.b.synchronize()
        return self.a.elapsed_time(self.b)


class BackProjector:
    """Projections (ntheta, nz, n) + angles prepared once for any number of back-projections (te_bp_prepare: the
    z-blocked re-layout and the cos / sin table live in one workspace tensor)."""

    def __init__(self, data, theta, workspace=None):
        _device.require_cuda()
        d = _f32_dev(data)
        if d.dim() != 3:
            raise ValueError("projections must be (ntheta, nz, n)")
        th = _f32_dev(theta).reshape(-1)
        self.ntheta, self.nz, self.n = (int(s) for s in d.shape)
        if th.numel() != self.ntheta:
            raise ValueError("theta has %d entries for %d projections" % (th.numel(), self.ntheta))
        L = _lib.lib()
        nbytes = int(L.te_bp_workspace_bytes(self.ntheta, self.nz, self.n))
        if workspace is not None and workspace.numel() * workspace.element_size() >= nbytes:
            self.ws = workspace
        else:
            self.ws = torch.empty(nbytes, dtype=torch.uint8, device=d.device)
        _lib.check(L.te_bp_prepare(d.data_ptr(), th.data_ptr(), self.ntheta, self.nz, self.n, self.ws.data_ptr(),
                                   _device.stream_ptr()), "te_bp_prepare")
        self._L = L

    def _dims(self):
        return self.ntheta, self.nz, self.n

    def box(self, center, stx, px, sty, py, stz, pz, out=None, masked=False, exact=False):
        if out is None:
            out = torch.empty((pz, py, px), dtype=torch.float32, device=self.ws.device)
        if out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous() or out.numel() != pz * py * px:
            raise TypeError("out must be a contiguous float32 CUDA tensor of %d voxels" % (pz * py * px))
        _lib.check(self._L.te_bp_box(self.ws.data_ptr(), *self._dims(), float(center), int(stx), int(px), int(sty), int(py),
                                     int(stz), int(pz), out.data_ptr(), 1 if masked else 0, 1 if exact else 0,
                                     _device.stream_ptr()), "te_bp_box")
        return out

    def volume(self, center, out=None, masked=False, exact=False):
        return self.box(center, 0, self.n, 0, self.n, 0, self.nz, out=out, masked=masked, exact=exact)

    def patches(self, center, points, wd, out_dtype=torch.float32, min_max=None, exact=False, out=None):
        """(npatch, wd, wd, wd) boxes at `points` (z, y, x); min_max fuses clip + (v - min) / (max - min)."""
        pts = _device.points_to_device(points) if not isinstance(points, torch.Tensor) else points
        n = int(pts.shape[0])
        if out is None:
            out = torch.empty((n, wd, wd, wd), dtype=out_dtype, device=self.ws.device)
        lo, hi = (0.0, 1.0) if min_max is None else (float(min_max[0]), float(min_max[1]))
        _lib.check(self._L.te_bp_patches(self.ws.data_ptr(), *self._dims(), float(center), pts.data_ptr(), n, int(wd),
                                         out.data_ptr(), _device.te_dtype(out), 0 if min_max is None else 1, lo, hi,
                                         1 if exact else 0, _device.stream_ptr()), "te_bp_patches")
        return out

    def points(self, center, pts, xy=False, exact=False):
        p = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.int32)).to(self.ws.device) if not isinstance(pts, torch.Tensor) \
            else pts.to(torch.int32).contiguous()
        if p.dim() != 2 or p.shape[1] != (2 if xy else 3):
            raise ValueError("pts must be (npts, %d)" % (2 if xy else 3))
        npts = int(p.shape[0])
        out = torch.zeros(npts * self.nz if xy else npts, dtype=torch.float32, device=self.ws.device)
        _lib.check(self._L.te_bp_points(self.ws.data_ptr(), *self._dims(), float(center), p.data_ptr(), npts, 1 if xy else 0,
                                        out.data_ptr(), 1 if exact else 0, _device.stream_ptr()), "te_bp_points")
        return out


def rec_pts(data, theta, center, pts, exact=False):
    """cuda_kernels.py:162-185: pts (npts, 3) z, y, x -> (npts,) float32."""
    t = _Timer()
    obj = BackProjector(data, theta).points(center, pts, xy=False, exact=exact)
    print("TIME rec_pts: %.2f ms" % t.stop())
    return _device.to_host_like(obj, data)


def rec_pts_xy(data, theta, center, pts, exact=False):
    """cuda_kernels.py:187-217: pts (npts, 2) sorted y, x -> flat (nz * npts) float32, index z * npts + i."""
    t = _Timer()
    obj = BackProjector(data, theta).points(center, pts, xy=True, exact=exact)
    print("TIME rec_pts: %.2f ms" % t.stop())
    return _device.to_host_like(obj, data)


def _in_place(obj, data, theta, center, masked, exact):
    t = _Timer()
    bp = BackProjector(data, theta)
    if _device.is_device_array(obj):
        if obj.dtype != torch.float32 or not obj.is_contiguous():
            raise TypeError("obj must be a contiguous float32 tensor")
        bp.volume(center, out=obj, masked=masked, exact=exact)
    else:
        d = _f32_dev(obj)
        bp.volume(center, out=d, masked=masked, exact=exact)
        obj[...] = d.cpu().numpy().reshape(obj.shape)
    return t.stop()


def rec_mask(obj, data, theta, center, exact=False):
    """cuda_kernels.py:219-242: voxels of obj (nz, n, n) that are > 0 are reconstructed in place, the others become 0;
    returns the elapsed milliseconds."""
    return _in_place(obj, data, theta, center, True, exact)


def rec_all(obj, data, theta, center, exact=False):
    """cuda_kernels.py:245-266: obj (nz, n, n) reconstructed in place; returns the elapsed milliseconds."""
    return _in_place(obj, data, theta, center, False, exact)


def rec_patch(data, theta, center, stx, px, sty, py, stz, pz, TIMEIT=False, exact=False):
    """cuda_kernels.py:269-302: sub-volume [stz:stz+pz, sty:sty+py, stx:stx+px] -> (pz, py, px) float32."""
    t = _Timer()
    obj = BackProjector(data, theta).box(center, stx, px, sty, py, stz, pz, exact=exact)
    t_gpu = t.stop()
    obj = _device.to_host_like(obj, data)
    return (obj, t_gpu) if TIMEIT else obj
