"""Where do 16-bit runs flip labels?  For the U-nets of the parity tests: voxel agreement with the torch-CPU oracle and
the distribution of |oracle logit| over the flipped voxels (prove the near-tie claim).

    python tools/margin_hist.py > profiles/r02_margin_hist.txt
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import nets_torch as O, phantom, weights as OW  # noqa: E402
from tomoencoders_b200 import Grid  # noqa: E402
import test_nets_gpu as T  # noqa: E402

M_A03W = dict(n_filters=[32, 64, 128], n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True] * 3, pool_size=[2, 2, 2])


def report(name, params, w, x, mm):
    for mode in ("fp16", "bf16"):
        labels, logits, ref = T._gpu_and_oracle_logits(params, w, x, mode, mm)
        flipped = (logits > 0) != (ref > 0)
        a = np.abs(ref)
        print("%s  %s: agreement %.6f  flipped %d of %d  logit std %.3f  max|gpu-oracle| %.4f" % (
            name, mode, 1.0 - flipped.mean(), int(flipped.sum()), flipped.size, ref.std(), np.abs(logits - ref).max()))
        if flipped.any():
            f = a[flipped]
            print("    |oracle logit| of the flipped voxels: median %.4f  p90 %.4f  p99 %.4f  max %.4f" % (
                np.median(f), np.percentile(f, 90), np.percentile(f, 99), f.max()))
        edges = [0, 0.005, 0.01, 0.02, 0.05, 0.1, 0.25, 0.5, 1.0, np.inf]
        print("    |logit| bin        voxels   flipped   flipped/voxels")
        for lo, hi in zip(edges[:-1], edges[1:]):
            m = (a >= lo) & (a < hi)
            print("    [%5.3f, %5.3f)  %9d  %8d   %.5f" % (lo, hi, int(m.sum()), int((flipped & m).sum()),
                                                          (flipped & m).sum() / max(1, m.sum())))


def main():
    vol, _ = phantom.volume_f16((128, 128, 128))
    x = Grid(vol.shape, width=64).extract(vol)[..., np.newaxis]
    mm = (float(vol[::4, ::4, ::4].min()), float(vol[::4, ::4, ::4].max()))
    report("M_a01 (8 x 64^3 phantom tiles)", T.M_A01, OW.draw(O.unet_spec(**T.M_A01), 11), x, mm)
    xr = np.random.default_rng(0).uniform(0, 1, (3, 32, 32, 32, 1)).astype(np.float32)
    report("M_a03-wide [32,64,128] (3 x 32^3 uniform noise)", M_A03W, OW.draw(O.unet_spec(**M_A03W), 31), xr, (0.0, 1.0))
    report("M_a03-wide [32,64,128] (8 x 64^3 phantom tiles)", M_A03W, OW.draw(O.unet_spec(**M_A03W), 31), x[:4], mm)


if __name__ == "__main__":
    main()
