"""Back-projection / FBP-filter throughput on one B200 (csrc/recon_ops.cu) through the host API on device tensors, CUDA
events on the launching stream, inputs larger than L2.  Unit: voxel-angle updates per second (one update = one
`f += g0 + (g1 - g0) * w / n`, cuda_kernels.py:118-120).  The kernel is fp32-pipe / issue bound: ~3 lane-ops per update
(FADD, FFMA, FADD) + the per-column address / weight arithmetic amortised over 16 z, against 148 SMs x 128 lanes x clock.
CPU leg: the reference's own kernels compiled for the host (oracle/_ref/librec_ref.so, one thread) on a bounded sample."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200 import Grid  # noqa: E402
from tomoencoders_b200 import reconstruction as RC  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    out = {}
    n, nz, ntheta = 2048, 32, 1500
    g = torch.Generator(device="cuda").manual_seed(0)
    data = torch.randn((ntheta, nz, n), device="cuda", generator=g)
    theta = torch.linspace(0, float(np.pi), ntheta + 1, device="cuda")[:-1].contiguous()
    center = n / 2 + 0.5
    out["shape"] = [ntheta, nz, n]
    ms = timed(lambda: RC.fbp_filter(data), 3)
    out["fbp_filter_ms"] = ms
    out["fbp_filter_GBs_rw"] = 2 * data.numel() * 4 / ms / 1e6
    ms = timed(lambda: RC.BackProjector(data, theta), 3)
    out["prepare_ms"] = ms
    bp = RC.BackProjector(data, theta)
    obj = torch.empty((nz, n, n), dtype=torch.float32, device="cuda")
    for exact in (False, True):
        ms = timed(lambda: bp.volume(center, out=obj, exact=exact), 3)
        # updates actually executed: voxels whose s0 is in range (inside the reconstruction circle all are)
        va = nz * n * n * ntheta
        out["rec_all_%s" % ("exact" if exact else "fast")] = {"ms": ms, "Gupdates_per_s": va / ms / 1e6}
    p = Grid((nz, n, n), width=32)
    for frac in (0.05, 0.25, 1.0):
        sel = p.select_random_sample(max(1, int(frac * len(p))))
        pts = torch.from_numpy(sel.points.astype(np.int32)).cuda()
        xo = torch.empty((len(sel), 32, 32, 32), dtype=torch.float16, device="cuda")
        ms = timed(lambda: bp.patches(center, pts, 32, min_max=(-50.0, 50.0), out=xo), 3)
        va = len(sel) * 32 ** 3 * ntheta
        out["patches_r%.2f" % frac] = {"n": len(sel), "ms": ms, "Gupdates_per_s": va / ms / 1e6}
        m = torch.empty((nz, n, n), dtype=torch.float32, device="cuda")

        def masked():
            RC.make_mask(m, sel.points, 32)
            bp.volume(center, out=m, masked=True)
        out["rec_mask_r%.2f_ms" % frac] = timed(masked, 3)
    clock = 1.9e9
    out["fp32_lane_ops_peak_per_s"] = 148 * 128 * clock
    out["fast_frac_of_3op_bound"] = out["rec_all_fast"]["Gupdates_per_s"] * 1e9 * 3 / (148 * 128 * clock)
    # end to end through the reference-facing call: HOST projections (4 z chunks), pinned double-buffered uploads
    # overlapped with filter + back-projection (+ segmentation), labels back on the host
    e_nz = 128
    host = np.random.default_rng(1).standard_normal((ntheta, e_nz, n)).astype(np.float32)
    th_h = np.linspace(0, np.pi, ntheta, endpoint=False).astype(np.float32)
    grid = Grid((e_nz, n, n), width=32)
    sel = grid.select_random_sample(len(grid) // 4)
    from tomoencoders_b200.neural_nets import SurfaceSegmenter
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="bench", mode="fp16", n_filters=[16, 32, 64],
                          n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True, True, True], pool_size=[2, 2, 2])
    host16 = np.clip(host * 1000 + 30000, 0, 65535).astype(np.uint16)       # raw detector counts: uploaded as uint16
    seg_kw = dict(segmenter=fe, segmenter_batch_size=256, rec_min_max=(-50.0, 50.0))
    for name, src, kw in (("recon_patches_3d_r0.25_host", host, {}), ("recon_segment_patches_3d_r0.25_host", host, seg_kw),
                          ("recon_segment_patches_3d_r0.25_host_uint16", host16, seg_kw)):
        RC.recon_patches_3d(src[:, :32], th_h, center, Grid((32, n, n), width=32).select_random_sample(64), **kw)   # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        x, _ = RC.recon_patches_3d(src, th_h, center, sel, **kw)
        dt = time.perf_counter() - t0
        out[name] = {"patches": len(sel), "s": dt, "voxels_per_s": len(sel) * 32 ** 3 / dt, "h2d_bytes": src.nbytes, "d2h_bytes": x.nbytes}
    try:
        from oracle import build_ref as R
        if R.built():
            nn, zz, tt = 256, 4, 180
            d = np.random.default_rng(0).standard_normal((tt, zz, nn)).astype(np.float32)
            th = np.linspace(0, np.pi, tt, endpoint=False).astype(np.float32)
            t0 = time.perf_counter()
            R.rec_all(d, th, nn / 2)
            dt = time.perf_counter() - t0
            out["cpu_reference_kernel"] = {"kind": "reference", "cores": 1, "sample": "rec_all %dx%dx%d" % (tt, zz, nn),
                                           "Gupdates_per_s": zz * nn * nn * tt / dt / 1e9, "s": dt}
    except Exception as e:  # noqa: BLE001
        out["cpu_reference_kernel"] = {"unavailable": str(e)}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
