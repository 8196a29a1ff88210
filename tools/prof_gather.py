"""Profiling target: random-corner uint8 gathers through the C ABI (cfg1: 4096 x 32^3 of 256^3; cfg3-like: 16384 x 64^3 of
512^3), index order and (z, y)-sorted processing order, three launches each.
    ncu --set full --clock-control none -k regex:gather_rows --launch-skip 2 -o gpurun_out/r02_gather python tools/prof_gather.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import _device, _lib  # noqa: E402


def case(vshape, P, n, seed, reps=3):
    L, st = _lib.lib(), _device.stream_ptr()
    v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device="cuda")
    rg = np.random.default_rng(seed)
    rp = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
    dp, dw = _device.points_to_device(rp), _device.points_to_device(np.full_like(rp, P))
    o8 = torch.empty((n, P, P, P), dtype=torch.uint8, device="cuda")
    perm = torch.argsort(dp[:, 0].to(torch.int64) * vshape[1] + dp[:, 1].to(torch.int64)).to(torch.int32)
    vs_, p_ = _lib.i64x3(vshape), _lib.i32x3((P,) * 3)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, fn in (("index order", lambda: L.te_gather(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                         _lib.TE_HINT_UNBINNED, st)),
                     ("sorted order", lambda: L.te_gather_perm(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                               _lib.TE_HINT_UNBINNED, perm.data_ptr(), st))):
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(fn())
            b.record()
            torch.cuda.synchronize()
            ts.append(1e3 * a.elapsed_time(b))
        alg = n * P ** 3 + min(n * P ** 3, v8.numel()) + 24 * n
        print("%s x %d^3 of %s, %s: %s us -> %.0f GB/s algorithmic" % (n, P, vshape, name, ["%.1f" % t for t in ts], alg / min(ts) / 1e3))


if __name__ == "__main__":
    case((256,) * 3, 32, 4096, 21)
    case((512,) * 3, 64, 16384, 22)
