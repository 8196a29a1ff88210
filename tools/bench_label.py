"""Times the labelling kernels (te_label = 7 launches, te_label_stats, te_tile_any) on synthetic masks (CUDA events) and
scipy.ndimage.label on a bounded sample on the host.  Algorithmic bytes of te_label: 1 B read + 4 B written per voxel."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200 import _device, _lib  # noqa: E402
from tomoencoders_b200.structures import voids as VV  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6545.6


def timed(fn, reps=5):
    fn()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def blobs(n, frac, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    v = torch.rand((1, 1, n // 8, n // 8, n // 8), device="cuda", generator=g)
    v = torch.nn.functional.interpolate(v, size=(n, n, n), mode="trilinear")[0, 0]
    return (v < torch.quantile(v.reshape(-1)[:: max(1, v.numel() // 1000000)], frac)).to(torch.uint8)


def main():
    L = _lib.lib()
    for n in (512, 1024):
        for name, vol in (("blobs 30 %", blobs(n, 0.3, 1)), ("noise 10 %", (torch.rand((n, n, n), device="cuda") < 0.1).to(torch.uint8))):
            shape = (n, n, n)
            nvox = n ** 3
            labels = torch.empty(shape, dtype=torch.int32, device="cuda")
            n_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
            ws = torch.empty(int(L.te_label_workspace_bytes(nvox)), dtype=torch.uint8, device="cuda")
            for conn in (26, 6):
                ms = timed(lambda: _lib.check(L.te_label(vol.data_ptr(), _lib.TE_U8, _lib.i64x3(shape), conn, labels.data_ptr(),
                                                         n_dev.data_ptr(), ws.data_ptr(), _device.stream_ptr()), "te_label"))
                nl = int(n_dev.item())
                print("te_label   %4d^3 %-10s conn %2d: %8.2f ms  %7.1f Mvox/ms... %6.1f GB/s algorithmic (%.1f %% of copy peak), %d components" %
                      (n, name, conn, ms, nvox / ms / 1e6, 5 * nvox / ms / 1e6, 100 * 5 * nvox / ms / 1e6 / PEAK, nl))
            bbox = torch.empty((max(nl, 1), 6), dtype=torch.int32, device="cuda")
            sizes = torch.empty(max(nl, 1), dtype=torch.int64, device="cuda")
            ms = timed(lambda: _lib.check(L.te_label_stats(labels.data_ptr(), _lib.i64x3(shape), nl, bbox.data_ptr(), sizes.data_ptr(),
                                                           _device.stream_ptr()), "te_label_stats"))
            print("te_label_stats %4d^3 %-10s: %8.2f ms  %6.1f GB/s (4 B/voxel read)" % (n, name, ms, 4 * nvox / ms / 1e6))
            flags = torch.zeros((n // 32) ** 3, dtype=torch.uint8, device="cuda")
            ms = timed(lambda: _lib.check(L.te_tile_any(vol.data_ptr(), 1, _lib.i64x3(shape), 32, flags.data_ptr(), _device.stream_ptr()),
                                          "te_tile_any"))
            print("te_tile_any    %4d^3 %-10s: %8.3f ms  %6.1f GB/s (1 B/voxel read)" % (n, name, ms, nvox / ms / 1e6))
            if n == 512:
                from scipy.ndimage import label as label_np
                h = vol[:256, :256, :256].cpu().numpy()
                t0 = time.perf_counter()
                label_np(h, structure=np.ones((3, 3, 3)))
                dt = time.perf_counter() - t0
                print("scipy.ndimage.label 256^3 %-10s conn 26 on the host (1 thread): %.2f s = %.1f Mvox/s" % (name, dt, h.size / dt / 1e6))
            del labels, ws


if __name__ == "__main__":
    main()
