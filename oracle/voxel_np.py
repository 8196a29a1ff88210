"""numpy restatement of the voxel pre-processing on the hot path.  TEST INFRASTRUCTURE.

Follows /root/reference/tomo_encoders/misc/voxel_processing.py:81-104 and the clip in
neural_nets/surface_segmenter.py:272-275.
"""
import numpy as np


def rescale_data(data, min_val, max_val):
    """voxel_processing.py:81-89: (data - min) / (max - min + 1e-12), numpy dtype promotion rules."""
    eps = 1e-12
    return (data - min_val) / (max_val - min_val + eps)


def find_min_max(vol, sampling_factor):
    """voxel_processing.py:91-104: min / max over vol[::s, ::s, ::s]."""
    ss = slice(None, None, sampling_factor)
    sub = vol[ss, ss, ss]
    return sub.min(), sub.max()


def clip_rescale_f32(x, min_max, clip):
    """The arithmetic the product kernels implement: everything in float32 (SURVEY.md H6 -- the
    reference's own result depends on the numpy promotion of the input dtype; the comparison is made
    at the latent / mask level with the tolerances of BASELINE.json.north_star)."""
    x = np.asarray(x).astype(np.float32)
    if min_max is None:
        return x
    lo, hi = np.float32(min_max[0]), np.float32(min_max[1])
    if clip:
        x = np.clip(x, lo, hi)
    return ((x - lo) / (hi - lo + np.float32(1e-12))).astype(np.float32)


# ---- normalisation / contrast / masking around the path -------------------
def modified_autocontrast(vol, s=0.01, normalize_sampling_factor=2):
    """voxel_processing.py:160-204: clamp values off the CDF of a 500-bin np.histogram of the
    sub-sampled volume (4-D: every patch, sub-sampled in space; 1-D: as is)."""
    sbin = slice(None, None, normalize_sampling_factor)
    if vol.ndim == 4:
        vals = vol[:, sbin, sbin, sbin].reshape(-1)
    elif vol.ndim == 3:
        vals = vol[sbin, sbin, sbin].reshape(-1)
    else:
        vals = vol
    slow, shigh = s if (type(s) == tuple and len(s) == 2) else (s, s)
    h, bins = np.histogram(vals, bins=500)
    c = np.cumsum(h)
    c_norm = c / np.max(c)
    return bins[np.argmin(np.abs(c_norm - slow))], bins[np.argmin(np.abs(c_norm - 1 + shigh))]


def histogram_by_edges(vals, edges):
    """The definition np.histogram's uniform-bin fast path converges to (numpy/lib/_histograms_impl.py:
    estimate, then one-step correction against the edges): counts[i] = #{edges[i] <= a < edges[i+1]},
    last bin closed.  The CUDA kernel implements exactly this; tests check it equals np.histogram."""
    vals = np.asarray(vals).reshape(-1).astype(edges.dtype)
    idx = np.searchsorted(edges, vals, side="right") - 1
    idx[vals == edges[-1]] = len(edges) - 2
    keep = (vals >= edges[0]) & (vals <= edges[-1])
    return np.bincount(idx[keep], minlength=len(edges) - 1).astype(np.int64)


def normalize_volume(vol, normalize_sampling_factor=1):
    """voxel_processing.py:207-253 (normalize_volume_gpu) without the chunked cupy staging: min / max
    of the sub-sampled volume, then _rescale_data on a float32 copy of every z-chunk, written back into
    ``vol`` (cast to vol.dtype).  Returns vol."""
    mn, mx = find_min_max(vol, normalize_sampling_factor)
    vol[...] = rescale_data(vol.astype(np.float32), mn, mx)
    return vol


def edge_map(Y):
    """voxel_processing.py:106-122: True where a voxel differs from a neighbour along any axis."""
    msk = np.zeros(Y.shape, dtype=np.uint8)
    for ax in range(3):
        a = [slice(None)] * 3
        b = [slice(None)] * 3
        a[ax], b[ax] = slice(None, -1), slice(1, None)
        d = Y[tuple(a)] != Y[tuple(b)]
        msk[tuple(a)][d] = 1
        msk[tuple(b)][d] = 1
    return msk > 0


def _cylinder(vol_shape, mask_fac):
    """voxel_processing.py:127-137 / :146-155: the uint8 cylinder indicator (inside = 1)."""
    assert vol_shape[1] == vol_shape[2], "must be a tomographic volume where shape y = shape x"
    n = vol_shape[1]
    rad = int(mask_fac * n / 2)
    pts = np.arange(-int(n // 2), int(np.ceil(n // 2)))
    yy, xx = np.meshgrid(pts, pts, indexing="ij")
    circ = (np.sqrt(yy ** 2 + xx ** 2) < rad).astype(np.uint8)
    return np.repeat(circ[np.newaxis, ...], vol_shape[0], axis=0)


def cylindrical_mask(out_vol, mask_fac, mask_val=0):
    """voxel_processing.py:124-140 (in place)."""
    out_vol[_cylinder(out_vol.shape, mask_fac) == 0] = mask_val


def get_values_cyl_mask(vol, mask_fac):
    """voxel_processing.py:142-157."""
    return vol[_cylinder(vol.shape, mask_fac) > 0]
