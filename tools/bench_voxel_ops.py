"""Kernel-only bandwidth of the normalisation / contrast / masking kernels (csrc/voxel_ops.cu) through the
host API on device tensors (CUDA events, L2 flushed).  Algorithmic bytes: histogram / min-max = bytes of
the rows the sub-sampled view touches at 32-byte sector granularity; rescale = read + write; edge_map =
read + 1 byte written per voxel; cylindrical_mask = bytes written outside the cylinder."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200.misc import voxel_processing as vp  # noqa: E402

PEAK = 6545.6
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:  # noqa: BLE001
    pass


def timed(fn, flush, reps=10):
    fn()
    torch.cuda.synchronize()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        flush.view(torch.int32)[: flush.numel() // 8].sum()   # read pass: the dirty lines of the write leave L2 before the timed call
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        tot += a.elapsed_time(b)
    return tot / reps


def line(name, ms, nbytes):
    print("%-58s %9.1f us  %7.1f GB/s  %5.1f%% of measured copy peak" % (name, ms * 1e3, nbytes / ms / 1e6, 100 * nbytes / ms / 1e6 / PEAK))


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    for dt in (torch.uint8, torch.float16, torch.float32):
        es = torch.empty((), dtype=dt).element_size()
        vol = (torch.rand((n, n, n), generator=g, device=dev) * 200).to(dt)
        tag = "%s %d^3" % (str(dt).replace("torch.", ""), n)
        for f in (1, 2):
            view = vol[::f, ::f, ::f]
            rows = view.shape[0] * view.shape[1]
            touched = rows * n * es if f * es <= 32 else view.numel() * 32      # sectors of the sampled rows
            line("%s min/max of [::%d]" % (tag, f), timed(lambda: vp._minmax_view(view), flush), touched)
            line("%s histogram(500) of [::%d] (incl. min/max, edges)" % (tag, f),
                 timed(lambda: vp._histogram_view(view, 500), flush), 2 * touched)
        if dt != torch.uint8:
            line("%s normalize_volume_gpu in place (min/max + rescale)" % tag,
                 timed(lambda: vp.normalize_volume_gpu(vol), flush), 3 * vol.numel() * es)
            vol = (torch.rand((n, n, n), generator=g, device=dev) * 200).to(dt)
        lab = (vol > 100).to(dt)
        line("%s edge_map" % tag, timed(lambda: vp.edge_map(lab), flush), vol.numel() * (es + 1))
        frac_out = 1.0 - 3.14159265 / 4.0
        line("%s cylindrical_mask(1.0)" % tag, timed(lambda: vp.cylindrical_mask(lab, 1.0, 0), flush), frac_out * vol.numel() * es)
        line("%s get_values_cyl_mask(1.0)" % tag, timed(lambda: vp.get_values_cyl_mask(lab, 1.0), flush),
             2 * (1 - frac_out) * vol.numel() * es)
        del vol, lab


if __name__ == "__main__":
    main()
