"""Generates tests/golden/voids_golden.npz by running the REFERENCE's own Voids class (structures/voids.py, imported
unmodified through oracle/ref_loader.py with scipy.ndimage.label standing in for cupyx.scipy.ndimage.label, as the
reference itself does on its CPU path) on seeded volumes.  Needs /root/reference: build container only.

    python tests/golden/make_voids_golden.py
"""
import os
import sys

import numpy as np
from scipy.ndimage import label as label_np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ref_loader, voids_np  # noqa: E402

ONES = np.ones((3, 3, 3), dtype=np.uint8)


def pack(v, prefix, out):
    out[prefix + "sizes"] = np.asarray(v["sizes"], dtype=np.int64)
    out[prefix + "cents"] = np.asarray(v["cents"], dtype=np.int64).reshape(-1, 3)
    out[prefix + "cpts"] = np.asarray(v["cpts"], dtype=np.int64).reshape(-1, 3)
    out[prefix + "boxes"] = np.asarray([[s[0].start, s[1].start, s[2].start, s[0].stop, s[1].stop, s[2].stop] for s in v["s_voids"]],
                                       dtype=np.int64).reshape(-1, 6)
    out[prefix + "masks"] = (np.concatenate([np.asarray(x, dtype=np.uint8).ravel() for x in v["x_voids"]])
                             if len(v["x_voids"]) else np.zeros(0, np.uint8))
    xb = v["x_boundary"] if "x_boundary" in v else []
    out[prefix + "boundary"] = np.asarray(xb, dtype=np.uint8)


def main():
    ref = ref_loader.load()
    out = {}
    # case A: label + count_voids + export_grid on a 48 x 40 x 56 phantom (26- and 6-connectivity), b = 2
    V = voids_np.void_phantom((48, 40, 56), 3)
    for conn, tag in ((26, "c26_"), (6, "c6_")):
        lab, n = label_np(V, structure=ONES if conn == 26 else None)
        out[tag + "labels"] = lab.astype(np.int32)
        out[tag + "n"] = np.int64(n)
    lab26 = out["c26_labels"]
    for dust, pad, tag in ((2, 0, "cv_d2_"), (0, 0, "cv_d0_"), (1, 2, "cv_d1p2_")):
        v = ref.RefVoids()
        v.pad_bb = pad
        v.count_voids(lab26.copy(), 2, dust, boundary_loc=(0, 0, 0))
        pack(v, tag, out)
        if len(v):
            p, r = v.export_grid(16)
            out[tag + "grid_points"], out[tag + "grid_r"] = np.asarray(p.points), np.float64(r)
            out[tag + "grid_wd"], out[tag + "grid_shape"] = np.int64(p.wd), np.asarray(p.vol_shape)
    # case B: import_from_grid.  Coarse voids from the phantom binned by 2, refined on the full-resolution mask.
    Vf = voids_np.void_phantom((64, 64, 96), 5, p_speck=0.003)
    Vb = Vf[::2, ::2, ::2].copy()
    labb, _ = label_np(Vb, structure=ONES)
    vb = ref.RefVoids().count_voids(labb, 2, 2)
    p_grid, _ = vb.export_grid(16)
    x_grid = p_grid.extract(Vf)
    vf = ref.RefVoids().import_from_grid(vb, x_grid, p_grid)
    pack(vb, "ig_b_", out)
    out["ig_grid_points"] = np.asarray(p_grid.points)
    pack(vf, "ig_f_", out)
    path = os.path.join(os.path.dirname(__file__), "voids_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", {k: v.shape for k, v in out.items() if v.ndim and v.size < 20})


if __name__ == "__main__":
    main()
