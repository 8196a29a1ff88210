"""Times te_tr_conv3_wgrad (tcgen05 weight gradient) on the layer shapes of M_a01 at batch 8 (CUDA events, L2 flushed)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200 import _device, _lib  # noqa: E402

SHAPES = [("dec0.conv2", 64, 32, 32), ("dec0.conv1a", 64, 64, 32), ("dec1.conv2", 32, 64, 64), ("dec1.conv1a", 32, 128, 64),
          ("dec2.conv2", 16, 128, 128), ("enc0.conv2", 64, 16, 32), ("bott.conv2", 8, 128, 256)]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = _device.device()
    L = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for name, s, ci, co in SHAPES:
        x = torch.randn((n, s, s, s, ci), device=dev).to(torch.bfloat16)
        dy = torch.randn((n, s, s, s, co), device=dev).to(torch.bfloat16)
        dw = torch.zeros((27, ci, co), device=dev)
        tx = _lib.TeTensor(x.data_ptr(), _lib.TE_BF16, ci, 0, ci, n, s, s, s)
        td = _lib.TeTensor(dy.data_ptr(), _lib.TE_BF16, co, 0, co, n, s, s, s)
        ts = []
        for it in range(6):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.check(L.te_tr_conv3_wgrad(ctypes.byref(tx), ctypes.byref(td), dw.data_ptr(), ci, 0, None, _device.stream_ptr()), "wgrad")
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b) * 1e3)
        us = sorted(ts[1:])[len(ts[1:]) // 2]
        fl = 2.0 * 27 * ci * co * n * s ** 3
        print("%-12s %3d^3 %3d -> %3d   %8.1f us   %7.1f TFLOP/s   (%.0f MB operands)" % (name, s, ci, co, us, fl / us / 1e6,
                                                                                          (x.numel() + dy.numel()) * 2 / 1e6))


if __name__ == "__main__":
    main()
