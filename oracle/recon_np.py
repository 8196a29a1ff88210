"""CPU oracle of the filtered back-projection row.  TEST INFRASTRUCTURE ONLY: imported by
tests/, tests/golden/make_recon_golden.py and tools/bench_recon.py's cpu_baseline leg, never by the product path.

numpy restatement, one IEEE fp32 rounding per reference operation, of
  * the five CuPy RawKernels of /root/reference/tomo_encoders/reconstruction/cuda_kernels.py:21-147
    (rec_pts :23-43, rec_pts_xy :47-68, rec :73-98, rec_mask :100-123, rec_all :125-147),
  * prep.py:21-31 calc_padding, :33-60 fbp_filter, :63-77 preprocess,
  * recon.py:369-381 make_mask, :237-253 extract_from_mask, :256-306 extract_segmented (the clip / rescale only),
    :135-233 recon_patches_3d (projection chunk -> filter -> mask -> rec_mask -> patches).

Parity status: PINNED for the back-projection kernels -- oracle/_ref/librec_ref.so is the reference's own kernel source
(the string literal of cuda_kernels.py, extracted at build time, compiled for the host by oracle/build_ref.py with a
thread-index shim, `__cosf/__sinf` -> cosf/sinf, -ffp-contract=off); tests/test_oracle_recon.py checks this restatement
against it and tests/golden/recon_golden.npz holds its outputs.  The GPU build of the reference (NVRTC) contracts FMAs and
uses the fast-math `__cosf/__sinf`, so it differs from either by rounding; that is what the FAST mode of the CUDA kernels
reproduces and why its tolerance is stated in the tests.  fbp_filter / preprocess run CuPy + cuFFT in the reference
(not runnable here): restated with numpy.fft in float32 -- parity unpinned for those two.
"""
import numpy as np

F = np.float32


def calc_padding(data_shape):
    """prep.py:21-31."""
    ntheta, nz, n = data_shape
    n_pad = n * (1 + 0.25 * 2)
    n_pad = int(np.ceil(n_pad / 8.0) * 8.0)
    pad_left = int((n_pad - n) // 2)
    pad_right = n_pad - n - pad_left
    return pad_left, pad_right


def fbp_filter(data):
    """prep.py:33-60, in place; float32 / complex64 like cupyx.scipy.fft on a float32 array."""
    pad_left, pad_right = calc_padding(data.shape)
    padded = np.pad(data, ((0, 0), (0, 0), (pad_left, pad_right)), mode="edge")
    t = np.fft.rfftfreq(padded.shape[2]).astype(F)
    spec = t * np.fft.rfft(padded, axis=2).astype(np.complex64)
    data[:] = np.fft.irfft(spec, n=padded.shape[2], axis=2).astype(F)[..., pad_left:-pad_right]
    return data


def preprocess(data, dark, flat):
    """prep.py:63-77, in place (median_filter with size [1,1,1] is the identity, so no voxel is replaced)."""
    data[:] = (data - dark) / np.maximum(flat - dark, F(1.0e-6))
    data[:] = -np.log(np.maximum(data, F(1.0e-6)))
    return data


def _trig(theta):
    t = np.asarray(theta, dtype=F).astype(np.float64)
    return np.cos(t).astype(F), np.sin(t).astype(F)


def _roundf(sp):
    """C roundf: half away from zero (np.round is half-to-even); exact in float64."""
    s = sp.astype(np.float64)
    return np.where(s >= 0, np.floor(s + 0.5), -np.floor(-s + 0.5)).astype(F)


def backproject(data, theta, center, xs, ys, zs=None):
    """The loop body shared by all five kernels for columns (xs[i], ys[i]) (int arrays of equal shape).

    zs None: all nz rows -> result shape (nz,) + xs.shape; else zs is an int array like xs (one z per point, rec_pts)
    -> result shape xs.shape.  cos / sin are evaluated in float64 and rounded to float32."""
    data = np.asarray(data, dtype=F)
    ntheta, nz, n = data.shape
    c, s = _trig(theta)
    xs = np.asarray(xs)
    ys = np.asarray(ys)
    fx = (xs.astype(np.int64) - n // 2).astype(F)
    fy = (ys.astype(np.int64) - n // 2).astype(F)
    center = F(center)
    nf = F(n)
    out = np.zeros(((nz,) if zs is None else ()) + xs.shape, F)
    for k in range(ntheta):
        sp = (fx * c[k] - fy * s[k]) + center                      # three roundings
        s0f = _roundf(sp)
        s0 = s0f.astype(np.int64)
        ok = (s0 >= 0) & (s0 < n - 1)
        s0c = np.where(ok, s0, 0)
        w = sp - s0f
        if zs is None:
            g0 = data[k][:, s0c]                                   # (nz,) + xs.shape
            g1 = data[k][:, s0c + 1]
        else:
            g0 = data[k][zs, s0c]
            g1 = data[k][zs, s0c + 1]
        t = g0 + ((g1 - g0) * w) / nf
        out = np.where(ok, out + t, out)
    return out


def rec_all(data, theta, center):
    """cuda_kernels.py:125-147 -> (nz, n, n)."""
    n = data.shape[2]
    ys, xs = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    return backproject(data, theta, center, xs, ys)


def rec_mask(obj, data, theta, center):
    """cuda_kernels.py:100-123: voxels with obj > 0 are reconstructed, all others become 0 (the store is unconditional)."""
    full = rec_all(data, theta, center)
    return np.where(obj > 0, full, F(0))


def rec_patch(data, theta, center, stx, px, sty, py, stz, pz):
    """cuda_kernels.py:73-98 -> (pz, py, px)."""
    ys, xs = np.meshgrid(sty + np.arange(py), stx + np.arange(px), indexing="ij")
    return backproject(data[:, stz:stz + pz], theta, center, xs, ys)


def rec_pts(data, theta, center, pts):
    """cuda_kernels.py:23-43: pts (npts, 3) z, y, x -> (npts,)."""
    pts = np.asarray(pts)
    return backproject(data, theta, center, pts[:, 2], pts[:, 1], zs=pts[:, 0])


def rec_pts_xy(data, theta, center, pts):
    """cuda_kernels.py:47-68: pts (npts, 2) y, x -> flat (nz * npts), index z * npts + i."""
    pts = np.asarray(pts)
    return backproject(data, theta, center, pts[:, 1], pts[:, 0]).reshape(-1)


def make_mask(shape, corner_pts, wd):
    """recon.py:369-381."""
    m = np.zeros(shape, F)
    for p in corner_pts:
        m[p[0]:p[0] + wd, p[1]:p[1] + wd, p[2]:p[2] + wd] = 1
    return m


def extract_from_mask(obj_mask, cpts, wd):
    """recon.py:237-253."""
    return [obj_mask[p[0]:p[0] + wd, p[1]:p[1] + wd, p[2]:p[2] + wd].copy() for p in cpts]


def normalize_batch(yp, rec_min_max):
    """recon.py:283-285."""
    lo, hi = F(rec_min_max[0]), F(rec_min_max[1])
    yp = np.clip(yp, lo, hi)
    return (yp - lo) / (hi - lo)


def recon_patches_3d(projs, theta, center, points, wd, apply_fbp=True, dark_flat=None):
    """recon.py:135-233 without a segmenter: returns (x (n, wd, wd, wd) float32, points re-ordered by z chunk).

    `points` (n, 3) z, y, x corners of a Grid of width wd; apply_fbp is accepted and ignored like in the reference
    (the filter always runs)."""
    points = np.asarray(points)
    n = projs.shape[2]
    x, pts_all = [], []
    for z_pt in np.unique(points[:, 0]):
        cpts = points[points[:, 0] == z_pt].copy()
        pts_all.append(cpts.copy())
        cpts[:, 0] = 0
        data = projs[:, z_pt:z_pt + wd, :].astype(F)
        if dark_flat is not None:
            preprocess(data, dark_flat[0][z_pt:z_pt + wd].astype(F), dark_flat[1][z_pt:z_pt + wd].astype(F))
        fbp_filter(data)
        mask = make_mask((wd, n, n), cpts, wd)
        obj = rec_mask(mask, data, theta, center)
        x.extend(extract_from_mask(obj, cpts, wd))
    return np.asarray(x, F), np.concatenate(pts_all, axis=0)
