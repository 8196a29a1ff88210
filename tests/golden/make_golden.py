"""Generates tests/golden/structures_golden.npz by running the REFERENCE's own Patches / Grid classes
(imported unmodified through oracle/ref_loader.py; needs /root/reference, so it only runs in the
build container).  The fixtures pin the oracle restatement (oracle/structures_np.py) and, through
it, the CUDA gather / scatter kernels.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ref_loader  # noqa: E402


def seeded_volume(shape, dtype, seed):
    g = np.random.default_rng(seed)
    if np.dtype(dtype) == np.uint8:
        return g.integers(0, 256, shape, dtype=np.uint8)
    if np.dtype(dtype) == np.uint16:
        return g.integers(0, 65536, shape, dtype=np.uint16)
    return g.standard_normal(shape).astype(dtype)


def main():
    ref = ref_loader.load()
    P, G = ref.RefPatches, ref.RefGrid
    out = {}
    vs = (351, 350, 340)
    # --- coordinate known answers (notebook outputs of the reference)
    out["mg2_points"] = P(vs, "multiple-grids", min_patch_size=(64,) * 3, max_stride=2, n_points=None).points
    p4 = P(vs, "multiple-grids", min_patch_size=(64,) * 3, max_stride=4, n_points=None)
    out["mg4_points"], out["mg4_widths"] = p4.points, p4.widths
    g2 = P(vs, "grid", patch_size=(64,) * 3, stride=2)
    out["grid_s2_points"], out["grid_s2_widths"] = g2.points, g2.widths
    rg = P(vs, "regular-grid", patch_size=(64,) * 3)
    out["rg_points"] = rg.points
    out["grid256_points"] = G((256,) * 3, width=32).points
    # --- gather / scatter on a small seeded volume, all dtypes
    svs = (40, 44, 48)
    for name, dt in (("u8", np.uint8), ("u16", np.uint16), ("f16", np.float16), ("f32", np.float32)):
        vol = seeded_volume(svs, dt, 100)
        # overlapping 'grid' patches of 32^3 + binned patches (width 64 -> 32, stride 2)
        pg = P(svs, "grid", patch_size=(16,) * 3, stride=1)
        out["gs_%s_grid_points" % name] = pg.points
        out["gs_%s_grid_extract" % name] = pg.extract(vol, (16,) * 3)
        pts = np.array([[0, 0, 0], [8, 12, 16], [3, 5, 7], [7, 11, 15], [24, 28, 32]])
        wds = np.array([[32] * 3, [32] * 3, [16] * 3, [32] * 3, [16] * 3])
        pb = P(svs, "data", points=pts, widths=wds)
        out["gs_%s_bin_points" % name], out["gs_%s_bin_widths" % name] = pb.points, pb.widths
        out["gs_%s_bin_extract" % name] = pb.extract(vol, (16,) * 3)
        # scatter with overlap: last index wins
        sub = seeded_volume((len(pg),) + (16,) * 3, dt, 101)
        vout = seeded_volume(svs, dt, 102)
        pg.fill_patches_in_volume(sub, vout)
        out["gs_%s_fill" % name] = vout
    # Grid extract / fill round trip
    gv = seeded_volume((32, 48, 32), np.uint8, 103)
    gg = G(gv.shape, width=16)
    out["grid_extract_u8"] = gg.extract(gv)
    # voxel_processing
    vv = seeded_volume((40, 44, 48), np.float32, 104)
    mn, mx = ref.ref_voxel._find_min_max(vv, 4)
    out["minmax_f32_s4"] = np.array([mn, mx], dtype=np.float64)
    out["rescale_f32"] = ref.ref_voxel._rescale_data(vv[:4, :4, :4], mn, mx)
    path = os.path.join(os.path.dirname(__file__), "structures_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
