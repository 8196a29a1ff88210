"""numpy restatement of the reference's Patches / Grid coordinate + gather + scatter semantics.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Every function cites the reference lines it follows
(paths relative to /root/reference/tomo_encoders/).  Pinned against the reference's own classes in
tests/test_oracle_pinned.py (when /root/reference is present) and against tests/golden fixtures.
"""
import numpy as np


# ----------------------------------------------------------------------------- coordinates
def check_stride(vol_shape, patch_size, stride):
    """structures/patches.py:216-232 (_check_stride)."""
    if stride is None:
        return tuple(patch_size)
    nd = len(vol_shape)
    max_possible = min(vol_shape[i] // patch_size[i] for i in range(nd))
    if stride > max_possible:
        raise ValueError("Cannot preserve aspect ratio with given combination of patch_size and stride.")
    return tuple(patch_size[i] * stride for i in range(nd))


def grid_points(vol_shape, patch_size, stride=1):
    """structures/patches.py:261-313 (_set_grid): overlapping grid covering the full volume.

    nsteps = ceil(m/p); step = (m-p)//(nsteps-1) (0 when m == p); z outermost, x innermost.
    """
    p = check_stride(vol_shape, patch_size, stride)
    m = list(vol_shape)
    nsteps = [int(np.ceil(m[i] / p[i])) for i in range(len(m))]
    step = [((m[i] - p[i]) // (nsteps[i] - 1)) if m[i] != p[i] else 0 for i in range(len(m))]
    axes = [np.arange(nsteps[i]) * step[i] for i in range(len(m))]
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(m))
    widths = np.tile(np.asarray(p), (len(pts), 1))
    return pts.astype(np.uint32), widths.astype(np.uint32)


def regular_grid_points(vol_shape, patch_size):
    """structures/patches.py:315-360 (_set_regular_grid) and structures/grid.py:64-105:
    nsteps = m//p, step = p, no overlap, volume cropped."""
    p = check_stride(vol_shape, patch_size, 1)
    m = list(vol_shape)
    nsteps = [int(m[i] // p[i]) for i in range(len(m))]
    axes = [np.arange(nsteps[i]) * p[i] for i in range(len(m))]
    pts = np.stack(np.meshgrid(*axes, indexing="ij"), axis=-1).reshape(-1, len(m))
    widths = np.tile(np.asarray(p), (len(pts), 1))
    return pts.astype(np.uint32), widths.astype(np.uint32)


def multiple_grids_points(vol_shape, min_patch_size, max_stride):
    """structures/patches.py:234-259 (_set_multiple_grids) with n_points=None."""
    pts, wds = [], []
    for stride in range(1, max_stride + 1):
        p, w = grid_points(vol_shape, min_patch_size, stride)
        pts.append(p)
        wds.append(w)
    return np.concatenate(pts, axis=0), np.concatenate(wds, axis=0)


def check_valid_points(points, widths, vol_shape):
    """structures/patches.py:416-428: invalid iff p < 0 or p + w > vol_shape (per axis)."""
    p = np.asarray(points).astype(np.int64)
    w = np.asarray(widths).astype(np.int64)
    bad = (p < 0) | (p + w > np.asarray(vol_shape, dtype=np.int64)[None, :])
    if bad.any():
        raise ValueError("Some points are invalid")


def calc_binning(widths, patch_size):
    """structures/patches.py:734-746 (_calc_binning)."""
    widths = np.asarray(widths)
    bin_vals = widths // np.asarray(patch_size)
    if np.sum(np.max(bin_vals, axis=1) != np.min(bin_vals, axis=1)) > 0:
        raise ValueError("aspect ratios of some patches don't match!! Cannot bin to patch_size")
    if np.any(bin_vals == 0):
        raise ValueError("patch_size cannot be larger than any given patch in the list")
    return bin_vals[:, 0]


# ----------------------------------------------------------------------------- gather / scatter
def extract(vol, points, widths, patch_size):
    """structures/patches.py:748-781 (extract) / structures/grid.py:313-336.

    out[i] = vol[z:z+wz:b, y:y+wy:b, x:x+wx:b] with b = widths//patch_size (isotropic); dtype kept.
    """
    points = np.asarray(points).astype(np.int64)
    widths = np.asarray(widths).astype(np.int64)
    b = calc_binning(widths, patch_size).astype(np.int64)
    out = []
    for i in range(len(points)):
        sl = tuple(slice(points[i, a], points[i, a] + widths[i, a], b[i]) for a in range(points.shape[1]))
        out.append(vol[sl])
    return np.asarray(out, dtype=vol.dtype)


def fill_patches_in_volume(sub_vols, points, widths, vol_out):
    """structures/patches.py:811-821 / structures/grid.py:338-348: sequential slice assignment,
    later index overwrites earlier on overlap, numpy assignment casts to vol_out.dtype."""
    points = np.asarray(points).astype(np.int64)
    widths = np.asarray(widths).astype(np.int64)
    for i in range(len(points)):
        sl = tuple(slice(points[i, a], points[i, a] + widths[i, a]) for a in range(points.shape[1]))
        vol_out[sl] = sub_vols[i]
