"""Kernel-only bandwidth of the gather / scatter / min-max kernels through the C ABI (CUDA events on
the launching stream, device-resident coordinates, no host work inside the timed region).

Algorithmic bytes: gather = n*P^3*s (writes) + min(n*P^3, |V|)*s (reads) +
n*24 (coordinates); scatter = n*P^3*(s_sub + s_vol)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200 import _device, _lib  # noqa: E402

PEAK = 6545.6
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timed(fn, reps=20, flush=None):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
            flush.view(torch.int32)[: flush.numel() // 8].sum()   # read pass: no dirty lines left for the timed call to evict
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps


def case(name, vshape, dtype, patch, pts, flush):
    es_ = torch.empty((), dtype=dtype).element_size()
    hint = _lib.TE_HINT_ALIGNED if not np.any((np.asarray(pts)[:, 2] * es_) % 16) else 0
    L = _lib.lib()
    dev = torch.device("cuda")
    n = len(pts)
    es = torch.empty((), dtype=dtype).element_size()
    vol = torch.randint(0, 255, vshape, device=dev, dtype=torch.uint8).to(dtype) if dtype != torch.uint8 else \
        torch.randint(0, 255, vshape, device=dev, dtype=torch.uint8)
    d_pts = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.uint32).view(np.int32)).to(dev)
    d_wds = torch.full((n, 3), patch, dtype=torch.int32, device=dev)
    out = torch.empty((n, patch, patch, patch), dtype=dtype, device=dev)
    st = _device.stream_ptr()

    def g():
        _lib.check(L.te_gather(vol.data_ptr(), es, _lib.i64x3(vshape), d_pts.data_ptr(), d_wds.data_ptr(), n,
                               _lib.i32x3((patch,) * 3), out.data_ptr(), hint, st))
    ms = timed(g, flush=flush)
    name = name + (" [tma]" if hint and es >= 2 else " [simt]")
    nv = n * patch ** 3
    alg = nv * es + min(nv, vol.numel()) * es + n * 24
    print("%-44s gather  %8.1f us  %7.1f GB/s  %5.1f%% of measured copy peak" % (name, ms * 1e3, alg / ms / 1e6, 100 * alg / ms / 1e6 / PEAK))
    # check against torch indexing on a few patches
    for i in (0, n // 2, n - 1):
        z, y, x = (int(v) for v in pts[i])
        assert torch.equal(out[i], vol[z:z + patch, y:y + patch, x:x + patch]), "gather mismatch"

    if dtype in (torch.uint8, torch.float16):
        act = torch.empty((n, patch, patch, patch), dtype=torch.float16, device=dev)

        def gn():
            _lib.check(L.te_gather_norm(vol.data_ptr(), _device.te_dtype(vol), _lib.i64x3(vshape), d_pts.data_ptr(),
                                        d_wds.data_ptr(), n, _lib.i32x3((patch,) * 3), 1, 1, 0.0, 255.0, act.data_ptr(),
                                        _lib.TE_F16, st))
        ms = timed(gn, flush=flush)
        alg = nv * 2 + min(nv, vol.numel()) * es + n * 24
        print("%-44s gather+norm->f16 %8.1f us  %7.1f GB/s  %5.1f%%" % (name, ms * 1e3, alg / ms / 1e6, 100 * alg / ms / 1e6 / PEAK))
    return vol, d_pts, out


def scatter_case(name, vshape, dtype, patch, pts, flush):
    L = _lib.lib()
    dev = torch.device("cuda")
    n = len(pts)
    es = torch.empty((), dtype=dtype).element_size()
    sub = torch.randint(0, 2, (n, patch, patch, patch), device=dev, dtype=torch.uint8).to(dtype)
    vol = torch.zeros(vshape, dtype=dtype, device=dev)
    d_pts = torch.from_numpy(np.ascontiguousarray(pts, dtype=np.uint32).view(np.int32)).to(dev)
    st = _device.stream_ptr()
    hint = _lib.TE_HINT_ALIGNED if not np.any((np.asarray(pts)[:, 2] * es) % 16) else 0

    def s():
        _lib.check(L.te_scatter(sub.data_ptr(), es, d_pts.data_ptr(), n, _lib.i32x3((patch,) * 3), vol.data_ptr(),
                                _lib.i64x3(vshape), None, hint, st))
    ms = timed(s, flush=flush)
    alg = 2 * n * patch ** 3 * es
    print("%-44s scatter %8.1f us  %7.1f GB/s  %5.1f%% of measured copy peak" % (name, ms * 1e3, alg / ms / 1e6, 100 * alg / ms / 1e6 / PEAK))
    i = n // 3
    z, y, x = (int(v) for v in pts[i])
    assert torch.equal(vol[z:z + patch, y:y + patch, x:x + patch], sub[i]), "scatter mismatch"


def grid_pts(vshape, p):
    zz, yy, xx = np.meshgrid(*[np.arange(0, s, p) for s in vshape], indexing="ij")
    return np.stack([zz.ravel(), yy.ravel(), xx.ravel()], axis=1)


def main():
    g = np.random.default_rng(3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")     # > 126 MB L2
    big = (1024, 1024, 1024) if "--big" in sys.argv else None
    case("cfg1 u8 256^3, 4096 random 32^3", (256,) * 3, torch.uint8, 32, g.integers(0, 256 - 32, (4096, 3)), flush)
    case("cfg2 f16 512^3, Grid 64^3 (512 tiles)", (512,) * 3, torch.float16, 64, grid_pts((512,) * 3, 64), flush)
    case("u8 512^3, Grid 64^3 (512 tiles)", (512,) * 3, torch.uint8, 64, grid_pts((512,) * 3, 64), flush)
    case("cfg3-like u8 512^3, 4096 random 64^3", (512,) * 3, torch.uint8, 64, g.integers(0, 512 - 64, (4096, 3)), flush)
    case("f32 512^3, Grid 64^3", (512,) * 3, torch.float32, 64, grid_pts((512,) * 3, 64), flush)
    if big:
        case("u8 1024^3, Grid 64^3 (4096 tiles)", big, torch.uint8, 64, grid_pts(big, 64), flush)
    scatter_case("cfg2 u8 mask 512^3, Grid 64^3", (512,) * 3, torch.uint8, 64, grid_pts((512,) * 3, 64), flush)
    scatter_case("f16 512^3, Grid 64^3", (512,) * 3, torch.float16, 64, grid_pts((512,) * 3, 64), flush)
    scatter_case("u8 256^3, Grid 32^3", (256,) * 3, torch.uint8, 32, grid_pts((256,) * 3, 32), flush)
    if big:
        scatter_case("u8 1024^3, Grid 64^3", big, torch.uint8, 64, grid_pts(big, 64), flush)


if __name__ == "__main__":
    main()
