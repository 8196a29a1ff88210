"""Reconstruction drivers: drop-in for /root/reference/tomo_encoders/reconstruction/recon.py (recon_all :22-62,
recon_all_gpu :64-94, recon_patches_3d :135-233, extract_from_mask :237-253, extract_segmented :256-306,
make_mask :369-381).

B200 design of recon_patches_3d: per z chunk the projections are uploaded, corrected and ramp-filtered on the device,
re-laid out once (te_bp_prepare), and the selected wd^3 patches are back-projected STRAIGHT into the patch array
(te_bp_patches) -- with a segmenter, already clipped / rescaled in the network's activation type, so the chunk's
tomogram, the float mask and the per-patch slicing copies of the reference never exist.  Results are the reference's:
rec_mask restricted to the patch boxes followed by slicing is the same arithmetic per voxel.
"""
import numpy as np
import torch

from .. import _device, _lib
from ..structures.grid import Grid
from .cuda_kernels import BackProjector, _f32_dev, _Timer
from .prep import _filter_dev, _preprocess_dev


def _chunk_to_device(projs, sl):
    """projs[:, sl, :] as a contiguous float32 CUDA tensor (host arrays: one strided copy + upload)."""
    if isinstance(projs, torch.Tensor):
        c = projs[:, sl, :]
        out = torch.empty(c.shape, dtype=torch.float32, device=_device.device())   # never aliases the caller's stack
        out.copy_(c)
        return out
    return torch.from_numpy(np.ascontiguousarray(projs[:, sl, :], dtype=np.float32)).to(_device.device())


def recon_all(projs, theta, center, nc, dark_flat=None, exact=False):
    """recon.py:22-62: full FBP in z chunks of nc rows; host in, host (nz, n, n) float32 out."""
    ntheta, nz, n = projs.shape
    obj_out = np.empty((nz, n, n), dtype=np.float32)
    th = _f32_dev(theta)
    for ic in range(int(np.ceil(nz / nc))):
        sl = slice(ic * nc, min(nz, (ic + 1) * nc))
        data = _chunk_to_device(projs, sl)
        if dark_flat is not None:
            _preprocess_dev(data, dark_flat[0][sl], dark_flat[1][sl])
        _filter_dev(data)
        obj = BackProjector(data, th).volume(center, exact=exact)
        obj_out[sl] = obj.cpu().numpy()
    return obj_out


def recon_all_gpu(projs, theta, center, obj, dark_flat=None, exact=False):
    """recon.py:64-94: whole projection stack on the device, obj (nz, n, n) float32 CUDA tensor filled in place."""
    if not _device.is_device_array(obj) or obj.dtype != torch.float32 or not obj.is_contiguous():
        raise TypeError("obj must be a contiguous float32 CUDA tensor")
    data = _chunk_to_device(projs, slice(None))
    if dark_flat is not None:
        _preprocess_dev(data, dark_flat[0], dark_flat[1])
    _filter_dev(data)
    BackProjector(data, _f32_dev(theta)).volume(center, out=obj, exact=exact)
    return


def make_mask(obj_mask, corner_pts, wd):
    """recon.py:369-381: obj_mask (float32 CUDA tensor) = 1 inside the wd^3 boxes, 0 elsewhere; returns milliseconds."""
    t = _Timer()
    if not _device.is_device_array(obj_mask) or obj_mask.dtype != torch.float32 or not obj_mask.is_contiguous():
        raise TypeError("obj_mask must be a contiguous float32 CUDA tensor")
    pts = np.ascontiguousarray(np.asarray(corner_pts).reshape(-1, 3), dtype=np.uint32)
    if len(pts) and (np.any(pts + wd > np.asarray(obj_mask.shape, dtype=np.int64))):
        raise ValueError("a box leaves the mask volume")
    d = _device.points_to_device(pts) if len(pts) else None
    _lib.check(_lib.lib().te_make_mask(obj_mask.data_ptr(), _lib.i64x3(obj_mask.shape), d.data_ptr() if len(pts) else None,
                                       len(pts), int(wd), _device.stream_ptr()), "te_make_mask")
    return t.stop()


def extract_from_mask(obj_mask, cpts, wd):
    """recon.py:237-253: list of host sub-volumes + milliseconds (one gather kernel + one download)."""
    t = _Timer()
    g = Grid(tuple(obj_mask.shape), initialize_by="data", points=np.asarray(cpts), width=wd)
    x = g.extract(obj_mask)
    x = x.cpu().numpy() if isinstance(x, torch.Tensor) else x
    return list(x), t.stop()


def _segment_normalized(segmenter, x_norm, batch_size):
    """Normalised patches (n, wd, wd, wd) in the activation type -> uint8 labels on the device."""
    model = segmenter.models["segmenter"]
    n, wd = int(x_norm.shape[0]), int(x_norm.shape[1])
    out = torch.empty((n, wd, wd, wd), dtype=torch.uint8, device=x_norm.device)
    for a in range(0, n, batch_size):
        b = min(n, a + batch_size)
        net = model.net_for((wd, wd, wd), min(batch_size, n))
        step = net.max_batch
        for a2 in range(a, b, step):
            b2 = min(b, a2 + step)
            net.unet_forward_into(x_norm[a2:b2], out[a2:b2])
    return out


def extract_segmented(obj_mask, cpts, wd, segmenter, batch_size, rec_min_max):
    """recon.py:256-306: patches of a reconstructed chunk -> clip / rescale -> segmenter -> uint8 (n, wd, wd, wd) host."""
    model = segmenter.models["segmenter"]
    net = model.net_for((wd, wd, wd), min(int(batch_size), max(1, len(cpts))))
    g = Grid(tuple(obj_mask.shape), initialize_by="data", points=np.asarray(cpts), width=wd)
    x = g.extract(obj_mask if _device.is_device_array(obj_mask) else _f32_dev(obj_mask))
    lo, hi = np.float32(rec_min_max[0]), np.float32(rec_min_max[1])
    x = ((torch.clamp(x, float(lo), float(hi)) - float(lo)) / float(hi - lo)).to(net.act_dtype).contiguous()
    return _segment_normalized(segmenter, x, int(batch_size)).cpu().numpy()


_COPY_THREADS = 8
_pool = None


def _copy_pool():
    global _pool
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor
        _pool = ThreadPoolExecutor(max_workers=_COPY_THREADS, thread_name_prefix="tomoenc-stage")
    return _pool


class _ChunkFeeder:
    """Projection chunks projs[:, z:z+nc, :] as float32 CUDA tensors, one chunk ahead of the consumer: the strided host
    copy goes into one of two pinned staging buffers and the upload runs on a side stream while the previous chunk is
    being filtered / back-projected (host arrays); device stacks are sliced on the device.  uint8 / uint16 / float16
    stacks (raw detector counts are uint16) are staged and uploaded in their own type -- half the PCIe bytes of the
    reference's host-side .astype(float32) -- and widened to float32 on the device."""

    def __init__(self, projs, z_list, nc):
        self.projs, self.z_list, self.nc = projs, [int(z) for z in z_list], int(nc)
        self.on_device = isinstance(projs, torch.Tensor) and projs.is_cuda
        self.i = 0
        if not self.on_device:
            dev = _device.device()
            shape = (int(projs.shape[0]), self.nc, int(projs.shape[2]))
            native = {"uint8": torch.uint8, "uint16": torch.uint16, "float16": torch.float16}
            name = str(projs.dtype).replace("torch.", "")
            self.sd = native.get(name, torch.float32)                   # staging / upload element type
            self.pinned = [torch.empty(shape, dtype=self.sd).pin_memory() for _ in range(2)]
            self.dev = [torch.empty(shape, dtype=self.sd, device=dev) for _ in range(2)]
            self.work = torch.empty(shape, dtype=torch.float32, device=dev) if self.sd != torch.float32 else None
            self.ready = [torch.cuda.Event(), torch.cuda.Event()]
            self.free = [torch.cuda.Event(), torch.cuda.Event()]       # the consumer is done with dev[b]
            self.staged = [torch.cuda.Event(), torch.cuda.Event()]     # the upload out of pinned[b] has finished
            self.stream = torch.cuda.Stream()
            self.used = [False, False]
            self._stage(0)

    def _stage(self, j):
        if j >= len(self.z_list):
            return
        b = j & 1
        z = self.z_list[j]
        if self.used[b]:
            self.staged[b].synchronize()                               # pinned[b] may be overwritten
        src = self.projs[:, z:z + self.nc, :]
        if isinstance(src, torch.Tensor):
            self.pinned[b].copy_(src)
        else:
            # strided host copy (one nc * n run per angle) split over a few threads: numpy releases the GIL in copyto
            dst = self.pinned[b].numpy()
            nth = dst.shape[0]
            step = -(-nth // _COPY_THREADS)
            list(_copy_pool().map(lambda a: np.copyto(dst[a:a + step], src[a:a + step], casting="unsafe"), range(0, nth, step)))
        with torch.cuda.stream(self.stream):
            if self.used[b]:
                self.stream.wait_event(self.free[b])
            self.dev[b].copy_(self.pinned[b], non_blocking=True)
            self.staged[b].record(self.stream)
            self.ready[b].record(self.stream)
        self.used[b] = True

    def next(self):
        """Chunk i (valid until the call after next()); chunk i + 1 is staged before returning."""
        j = self.i
        self.i += 1
        if self.on_device:
            z = self.z_list[j]
            out = torch.empty((self.projs.shape[0], self.nc, self.projs.shape[2]), dtype=torch.float32, device=self.projs.device)
            out.copy_(self.projs[:, z:z + self.nc, :])
            return out
        b = j & 1
        torch.cuda.current_stream().wait_event(self.ready[b])
        if self.work is None:
            return self.dev[b]
        d = self.dev[b]                                                # exact widening: (x - 0) / 1 in fp32
        _lib.check(_lib.lib().te_rescale(d.data_ptr(), _device.te_dtype(d), self.work.data_ptr(), _lib.TE_F32, d.numel(), 0.0, 1.0,
                                         _device.stream_ptr()), "te_rescale")
        return self.work

    def release(self, j):
        """The consumer has enqueued all its work on chunk j: stage chunk j + 1 (overlaps with that work)."""
        if self.on_device:
            return
        self.free[j & 1].record(torch.cuda.current_stream())
        self._stage(j + 1)


def recon_patches_3d(projs, theta, center, p3d, apply_fbp=True, TIMEIT=False, segmenter=None, segmenter_batch_size=256,
                     dark_flat=None, rec_min_max=None, exact=False):
    """recon.py:135-233: reconstruct (and optionally segment) the patches of Grid `p3d` chunk by chunk along z.

    Returns (x, p3d_new[, times]): x (n, wd, wd, wd) float32 reconstructions, or uint8 labels with a segmenter; p3d_new
    lists the patches in the order of x (grouped by z like the reference).  times rows: [ntheta, nc, n, t_cpu2gpu,
    t_filt, t_mask, t_rec(, t_gpu2cpu)] in milliseconds (t_mask is 0: no mask is built here; with TIMEIT the chunks are
    timed one by one, which serialises the upload / compute overlap).  The upload of chunk i + 1 overlaps the filter,
    back-projection and segmentation of chunk i; results stay on the device until the last chunk."""
    _device.require_cuda()
    if segmenter is not None and rec_min_max is None:
        raise ValueError("rec_min_max is required with a segmenter")
    wd = int(p3d.wd)
    points = np.asarray(p3d.points)
    ntheta, nc, n = projs.shape[0], wd, projs.shape[2]
    th = _f32_dev(theta)
    act_dtype = torch.float32
    if segmenter is not None:
        act_dtype = segmenter.models["segmenter"].net_for((wd,) * 3, min(int(segmenter_batch_size), len(points))).act_dtype
    z_pts = np.unique(points[:, 0])
    feeder = _ChunkFeeder(projs, z_pts, nc)
    x, times, cpts_all = [], [], []
    ws = None
    for j, z_pt in enumerate(z_pts):
        cpts = points[points[:, 0] == z_pt]
        cpts_all.append(cpts.copy())
        local = cpts.copy()
        local[:, 0] = 0
        t = _Timer() if TIMEIT else None
        data = feeder.next()
        if dark_flat is not None:
            _preprocess_dev(data, dark_flat[0][z_pt:z_pt + nc], dark_flat[1][z_pt:z_pt + nc])
        t_cpu2gpu = t.stop() if TIMEIT else 0.0
        t = _Timer() if TIMEIT else None
        _filter_dev(data)
        t_filt = t.stop() if TIMEIT else 0.0
        t = _Timer() if TIMEIT else None
        bp = BackProjector(data, th, workspace=ws)
        ws = bp.ws
        if segmenter is not None:
            xn = bp.patches(center, local, wd, out_dtype=act_dtype, min_max=rec_min_max, exact=exact)
            t_rec = t.stop() if TIMEIT else 0.0
            x.append(_segment_normalized(segmenter, xn, int(segmenter_batch_size)))
            times.append([ntheta, nc, n, t_cpu2gpu, t_filt, 0.0, t_rec])
        else:
            x.append(bp.patches(center, local, wd, exact=exact))
            t_rec = t.stop() if TIMEIT else 0.0
            times.append([ntheta, nc, n, t_cpu2gpu, t_filt, 0.0, t_rec, 0.0])
        feeder.release(j)                         # all device work of chunk j is enqueued: stage chunk j + 1 behind it
    cpts_all = np.concatenate(cpts_all, axis=0)
    t = _Timer()
    x = torch.cat(x, dim=0).cpu().numpy()
    t_gpu2cpu = t.stop()
    if segmenter is None and times:
        times[-1][-1] = t_gpu2cpu                 # one download at the end
    p3d = Grid(p3d.vol_shape, initialize_by="data", points=cpts_all, width=wd)
    if TIMEIT:
        return x, p3d, np.asarray(times)
    return x, p3d
