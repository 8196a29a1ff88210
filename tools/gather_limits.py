"""What bounds the random-corner uint8 gather?  Times (CUDA events, L2 flushed) on the cfg3-like case
(16 384 x 64^3 patches of a 512^3 uint8 volume) and the cfg1 case (4096 x 32^3 of 256^3):
  memset of the patch array (write-only stream), a device-to-device copy of the same bytes,
  the row-group gather with x corners forced to multiples of 128 / 32 / 16 / 4 / 1 bytes (sector over-fetch of the
  unaligned rows), with all corners equal (volume side served by a few L2 lines), in index and (z, y)-sorted order.
    python tools/gather_limits.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import _device, _lib  # noqa: E402


def timed_us(fn, flush, flush_r, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        flush_r.sum()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(1e3 * a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def case(vshape, P, n, seed):
    L, st = _lib.lib(), _device.stream_ptr()
    v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device="cuda")
    o8 = torch.empty((n, P, P, P), dtype=torch.uint8, device="cuda")
    src = torch.empty_like(o8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_r = torch.zeros(256 << 18, dtype=torch.int32, device="cuda")
    vs_, p_ = _lib.i64x3(vshape), _lib.i32x3((P,) * 3)
    nbytes = o8.numel()
    print("== %d x %d^3 of %s uint8: %.1f MB written" % (n, P, vshape, nbytes / 1e6))
    t, ta = timed_us(lambda: o8.zero_(), flush, flush_r)
    print("  memset                       %8.1f us (avg %8.1f)  %6.0f GB/s written" % (t, ta, nbytes / t / 1e3))
    t, ta = timed_us(lambda: o8.copy_(src), flush, flush_r)
    print("  d2d copy                     %8.1f us (avg %8.1f)  %6.0f GB/s read+written" % (t, ta, 2 * nbytes / t / 1e3))
    rg = np.random.default_rng(seed)
    rp0 = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
    for name, align, same in (("x % 128 == 0", 128, False), ("x % 32 == 0", 32, False), ("x % 16 == 0", 16, False),
                              ("x % 4 == 0", 4, False), ("random", 1, False), ("all corners (3, 5, 7)", 1, True)):
        rp = rp0.copy()
        rp[:, 2] = rp[:, 2] // align * align
        if same:
            rp[:] = (3, 5, 7)
        dp, dw = _device.points_to_device(rp), _device.points_to_device(np.full_like(rp, P))
        perm = torch.argsort(dp[:, 0].to(torch.int64) * vshape[1] + dp[:, 1].to(torch.int64)).to(torch.int32)
        t0, _ = timed_us(lambda: _lib.check(L.te_gather(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                        _lib.TE_HINT_UNBINNED, st)), flush, flush_r)
        t1, _ = timed_us(lambda: _lib.check(L.te_gather_perm(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                             _lib.TE_HINT_UNBINNED, perm.data_ptr(), st)), flush, flush_r)
        print("  gather %-22s index order %8.1f us, sorted %8.1f us  (%6.0f GB/s written)" % (name, t0, t1, nbytes / t1 / 1e3))


if __name__ == "__main__":
    case((512,) * 3, 64, 16384, 22)
    case((256,) * 3, 32, 4096, 21)
