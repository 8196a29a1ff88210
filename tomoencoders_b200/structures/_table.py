"""Data-plane calls shared by Patches and Grid: gather (extract) and scatter (fill) on the GPU.

The coordinate bookkeeping stays in numpy on the host (it is O(n) metadata); the voxel movement is
done by libtomoenc.so (csrc/gather_scatter.cu) through the C ABI in include/tomoenc.h.
"""
import numpy as np
import torch

from .. import _device, _lib


def as_3d(vol_shape):
    """2-D volumes (images) are handled as a single-slice 3-D volume."""
    vs = tuple(int(v) for v in vol_shape)
    if len(vs) == 2:
        return (1,) + vs
    if len(vs) != 3:
        raise ValueError("vol_shape must have 2 or 3 dimensions")
    return vs


def pad_cols(a, fill):
    """(n,2) -> (n,3) by prepending a z column (2-D support)."""
    a = np.asarray(a)
    if a.shape[1] == 3:
        return a
    return np.concatenate([np.full((len(a), 1), fill, dtype=a.dtype), a], axis=1)


def gather(vol, vol_shape, points, widths, patch_size):
    """out[i] = vol[z:z+wz:b, y:y+wy:b, x:x+wx:b]; returns (n,) + patch_size, dtype of vol.

    vol: numpy array, torch CPU tensor or torch CUDA tensor; the result lives where vol lives.
    """
    ndim = len(vol_shape)
    vs3 = as_3d(vol_shape)
    n = len(points)
    t_vol = _device.to_device(vol)
    if tuple(t_vol.shape) != tuple(vol_shape):
        raise AssertionError("Shape of big volume does not match vol_shape attribute of patches data")
    es = t_vol.element_size()
    if es == 8:
        # float64 / int64 (numpy's defaults): data movement only, so an unbinned gather is the same gather over pairs of
        # 32-bit words (x coordinates, widths and patch extent doubled); a decimating gather would need a 64-bit element
        # path and is refused up front
        if n and np.any(np.asarray(widths, dtype=np.int64) != np.asarray(patch_size, dtype=np.int64)[None, :]):
            raise TypeError("binned extract of a %s volume is not supported: cast the volume to float32" % (t_vol.dtype,))
        dbl = np.ones(ndim, dtype=np.int64)
        dbl[-1] = 2
        out32 = gather(t_vol.view(torch.int32), tuple(int(v) * int(d) for v, d in zip(vol_shape, dbl)),
                       np.asarray(points, dtype=np.int64) * dbl[None, :], np.asarray(widths, dtype=np.int64) * dbl[None, :],
                       tuple(int(p) * int(d) for p, d in zip(patch_size, dbl)))
        return _device.to_host_like(out32.view(t_vol.dtype), vol)
    if es not in (1, 2, 4):
        raise TypeError("unsupported volume dtype %s" % (t_vol.dtype,))
    patch3 = (1,) * (3 - ndim) + tuple(int(p) for p in patch_size)
    out = torch.empty((n,) + patch3, dtype=t_vol.dtype, device=t_vol.device)
    if n > 0:
        pts3 = pad_cols(points, 0)
        wds3 = np.asarray(widths)
        if wds3.shape[1] == 2:
            # the kernels derive the (isotropic) binning from the z column, widths[i][0] // patch[0]: a 2-D table is a
            # single-slice volume whose padded z width must carry the binning of the image axes (a z extent of b
            # sampled with step b is still exactly one slice)
            binning = (wds3[:, -1].astype(np.int64) // int(patch_size[-1])).astype(wds3.dtype)
            wds3 = np.concatenate([binning[:, None], wds3], axis=1)
        d_pts = _device.points_to_device(pts3)
        d_wds = _device.points_to_device(wds3)
        L = _lib.lib()
        hint = _aligned_hint(pts3, wds3, patch3, es)
        gathered = n * int(np.prod(patch3))
        if n >= 1024 and gathered >= 4 * int(np.prod(vs3)) and not (hint & _lib.TE_HINT_ALIGNED):
            # many overlapping patches at scattered corners: process them in (z, y) order so that the slab the running
            # blocks read stays in L2 (the output order does not change)
            key = d_pts[:, 0].to(torch.int64) * int(vs3[1]) + d_pts[:, 1].to(torch.int64)
            perm = torch.argsort(key).to(torch.int32)
            _lib.check(L.te_gather_perm(t_vol.data_ptr(), es, _lib.i64x3(vs3), d_pts.data_ptr(), d_wds.data_ptr(), n,
                                        _lib.i32x3(patch3), out.data_ptr(), hint, perm.data_ptr(), _device.stream_ptr()),
                       "te_gather_perm")
        else:
            _lib.check(L.te_gather(t_vol.data_ptr(), es, _lib.i64x3(vs3), d_pts.data_ptr(), d_wds.data_ptr(), n,
                                   _lib.i32x3(patch3), out.data_ptr(), hint, _device.stream_ptr()), "te_gather")
    out = out.reshape((n,) + tuple(int(p) for p in patch_size))
    return _device.to_host_like(out, vol)


_DISJOINT_CACHE = {}
_HINT_CACHE = {}


def _aligned_hint(points3, widths3, patch3, es):
    """TE_HINT_ALIGNED when every patch is unbinned and starts on a 16-byte boundary in x (every
    Grid / regular-grid tiling whose width * itemsize is a multiple of 16): libtomoenc may then use
    its TMA-staged kernels (include/tomoenc.h).  Memoised on the table contents for small tables."""
    pts = np.ascontiguousarray(points3)
    wds = None if widths3 is None else np.ascontiguousarray(widths3)
    if len(pts) > 65536:
        return _aligned_hint_uncached(pts, wds, patch3, es)
    key = (pts.shape, pts.tobytes(), None if wds is None else wds.tobytes(), tuple(patch3), es)
    r = _HINT_CACHE.get(key)
    if r is None:
        if len(_HINT_CACHE) >= 16:
            _HINT_CACHE.pop(next(iter(_HINT_CACHE)))
        r = _HINT_CACHE[key] = _aligned_hint_uncached(pts, wds, patch3, es)
    return r


def _aligned_hint_uncached(pts, wds, patch3, es):
    if wds is not None and np.any(np.asarray(wds, dtype=np.int64) != np.asarray(patch3, dtype=np.int64)[None, :]):
        return 0
    if np.any((pts[:, 2].astype(np.int64) * es) % 16):
        return _lib.TE_HINT_UNBINNED
    return _lib.TE_HINT_ALIGNED | _lib.TE_HINT_UNBINNED


def _disjoint(points, width, vs3):
    """True when the patches are provably pairwise disjoint: equal widths, corners on the width
    lattice, no duplicates (always the case for Grid / regular-grid).  Memoised on the table
    contents for small tables (the same tiling is stitched call after call)."""
    pts = np.ascontiguousarray(points)
    if len(pts) > 65536:
        return _disjoint_uncached(pts, width, vs3)
    key = (pts.shape, str(pts.dtype), pts.tobytes(), tuple(int(w) for w in width), tuple(int(v) for v in vs3))
    r = _DISJOINT_CACHE.get(key)
    if r is None:
        if len(_DISJOINT_CACHE) >= 16:
            _DISJOINT_CACHE.pop(next(iter(_DISJOINT_CACHE)))
        r = _DISJOINT_CACHE[key] = _disjoint_uncached(pts, width, vs3)
    return r


def _disjoint_uncached(points, width, vs3):
    p = np.asarray(points, dtype=np.int64)
    w = np.asarray(width, dtype=np.int64)
    if np.any(p % w[None, :]):
        return False
    cells = p // w[None, :]
    dims = np.asarray(vs3, dtype=np.int64) // w + 1
    lin = (cells[:, 0] * dims[1] + cells[:, 1]) * dims[2] + cells[:, 2]
    return len(np.unique(lin)) == len(lin)


def scatter(sub_vols, vol_out, vol_shape, points, widths):
    """Sequential-assignment semantics of fill_patches_in_volume: vol_out[slices_i] = sub_vols[i]
    for i = 0..n-1, later patches overwrite earlier ones; values are cast to vol_out.dtype the way a
    numpy assignment would.  vol_out is modified in place (numpy or torch, host or device)."""
    ndim = len(vol_shape)
    vs3 = as_3d(vol_shape)
    n = len(points)
    if len(sub_vols) != n:
        raise AssertionError("number of sub-volumes do not match the number of items in patches")
    if tuple(vol_out.shape) != tuple(vol_shape):
        raise AssertionError("shape of volume enclosing the patches does not match that of the output volume")
    if n == 0:
        return
    widths = np.asarray(widths)
    # group by width (Grid and fixed-width Patches have a single group); groups are processed in
    # order of their first patch index only when they are provably independent, otherwise the
    # whole call goes through the ordered (owner-map) path group by group in index order.
    uniq = np.unique(widths, axis=0)
    if len(uniq) != 1:
        # variable-width patches: order matters across groups, fall back to index-ordered runs
        runs = _runs_of_equal_width(widths)
    else:
        runs = [(0, n)]
    host_out = not _device.is_device_array(vol_out)
    t_out = _device.to_device(vol_out)
    es = t_out.element_size()
    if es == 8:
        # 64-bit volumes: the same scatter over pairs of 32-bit words (see gather)
        dbl = np.ones(ndim, dtype=np.int64)
        dbl[-1] = 2
        if isinstance(sub_vols, (list, tuple)):
            subs32 = [_numpy_cast(_device.to_device(e), t_out.dtype).contiguous().view(torch.int32) for e in sub_vols]
        else:
            subs32 = _numpy_cast(_device.to_device(sub_vols), t_out.dtype).contiguous().view(torch.int32)
        scatter(subs32, t_out.view(torch.int32), tuple(int(v) * int(d) for v, d in zip(vol_shape, dbl)),
                np.asarray(points, dtype=np.int64) * dbl[None, :], widths.astype(np.int64) * dbl[None, :])
        _copy_back(t_out, vol_out, host_out)
        return
    if es not in (1, 2, 4):
        raise TypeError("unsupported volume dtype %s" % (t_out.dtype,))
    L = _lib.lib()
    owner = None
    for (a, b) in runs:
        w3 = (1,) * (3 - ndim) + tuple(int(v) for v in widths[a])
        sv = sub_vols[a:b]
        if isinstance(sv, (list, tuple)):
            # a list of per-patch arrays (what extract(vol, patch_size=None) returns): torch tensors stay where they
            # are (a list of CUDA tensors cannot go through numpy), anything else is stacked on the host
            if any(isinstance(e, torch.Tensor) for e in sv):
                dev_of = next((e.device for e in sv if isinstance(e, torch.Tensor) and e.is_cuda), None)
                sv = torch.stack([(e if isinstance(e, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(e))).to(
                    dev_of if dev_of is not None else "cpu") for e in sv])
            else:
                sv = np.asarray(sv)
        t_sub = _device.to_device(sv)
        if tuple(t_sub.shape[1:]) != tuple(int(v) for v in widths[a]):
            raise AssertionError("width mismatch between sub_vol and patch")
        if t_sub.dtype != t_out.dtype:
            t_sub = _numpy_cast(t_sub, t_out.dtype)
        pts = pad_cols(np.asarray(points[a:b]), 0)
        d_pts = _device.points_to_device(pts)
        row_ok = (w3[2] * es) % 16 == 0
        if row_ok and _disjoint(pts, w3, vs3):
            owner_ptr = None
        else:
            if owner is None:
                owner = torch.empty(vs3, dtype=torch.int32, device=t_out.device)
            owner_ptr = owner.data_ptr()
        hint = 0 if owner_ptr is not None else _aligned_hint(pts, None, w3, es)
        _lib.check(L.te_scatter(t_sub.data_ptr(), es, d_pts.data_ptr(), b - a, _lib.i32x3(w3), t_out.data_ptr(),
                                _lib.i64x3(vs3), owner_ptr, hint, _device.stream_ptr()), "te_scatter")
    _copy_back(t_out, vol_out, host_out)


def _copy_back(t_out, vol_out, host_out):
    if host_out:
        res = t_out.cpu()
        if isinstance(vol_out, torch.Tensor):
            vol_out.copy_(res)
        else:
            vol_out[...] = res.numpy()
    elif t_out.data_ptr() != vol_out.data_ptr():
        vol_out.copy_(t_out)


def _runs_of_equal_width(widths):
    change = np.any(widths[1:] != widths[:-1], axis=1)
    starts = np.concatenate([[0], np.nonzero(change)[0] + 1])
    ends = np.concatenate([starts[1:], [len(widths)]])
    return [(int(a), int(b)) for a, b in zip(starts, ends)]


def _numpy_cast(t, dtype):
    """numpy's assignment cast (C-style truncation for float -> integer)."""
    if dtype in (torch.uint8, torch.uint16) and t.dtype.is_floating_point:
        return t.to(torch.int64).to(dtype)
    return t.to(dtype)
