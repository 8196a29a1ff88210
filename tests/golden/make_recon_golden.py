"""Generates tests/golden/recon_golden.npz with the reference's OWN back-projection kernels: the CUDA C source string of
/root/reference/tomo_encoders/reconstruction/cuda_kernels.py compiled for the host by oracle/build_ref.py (run in the
container where /root/reference is mounted; the fixture is committed, the .so is not).

    python tests/golden/make_recon_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from oracle import build_ref as R  # noqa: E402


def case_inputs(seed, ntheta, nz, n):
    """Seeded projections: smooth sinusoids + noise, float32 (ntheta, nz, n); angles over [0, pi)."""
    g = np.random.default_rng(seed)
    s = np.arange(n, dtype=np.float64)
    k = np.arange(ntheta, dtype=np.float64)[:, None, None]
    z = np.arange(nz, dtype=np.float64)[None, :, None]
    data = np.sin(0.21 * s + 0.05 * k + 0.3 * z) + 0.5 * np.cos(0.043 * s * (1 + 0.01 * z)) + 0.1 * g.standard_normal((ntheta, nz, n))
    theta = np.linspace(0, np.pi, ntheta, endpoint=False)
    return data.astype(np.float32), theta.astype(np.float32)


CASES = {
    # name: (seed, ntheta, nz, n, center)
    "a": (21, 45, 20, 72, 36.5),     # nz not a multiple of the 16-row z block, half-integer centre (exact .5 ties)
    "b": (22, 30, 16, 64, 31.0),     # power-of-two n
    "c": (23, 33, 37, 40, 17.3),     # three z blocks with a ragged tail, off-centre axis
}


def main():
    out = {}
    for name, (seed, ntheta, nz, n, center) in CASES.items():
        data, theta = case_inputs(seed, ntheta, nz, n)
        g = np.random.default_rng(seed + 100)
        out[name + "_rec_all"] = R.rec_all(data, theta, center)
        mask = (g.uniform(size=(nz, n, n)) > 0.6).astype(np.float32) * g.uniform(0.5, 2.0, (nz, n, n)).astype(np.float32)
        mask[g.uniform(size=mask.shape) > 0.95] = -1.0
        out[name + "_mask"] = mask
        out[name + "_rec_mask"] = R.rec_mask(mask, data, theta, center)
        box = np.array([5, n - 17, 3, n - 9, 3, nz - 4], np.int32)    # stx, px, sty, py, stz, pz
        out[name + "_box"] = box
        out[name + "_rec_patch"] = R.rec_patch(data, theta, center, *box)
        pts = np.stack([g.integers(0, nz, 257), g.integers(0, n, 257), g.integers(0, n, 257)], 1).astype(np.int32)
        out[name + "_pts"] = pts
        out[name + "_rec_pts"] = R.rec_pts(data, theta, center, pts)
        out[name + "_rec_pts_xy"] = R.rec_pts_xy(data, theta, center, pts[:, 1:])
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "recon_golden.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
