"""One te_label of a 512^3 blobs mask per connectivity (for `ncu --metrics gpu__time_duration.sum`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from bench_label import blobs  # noqa: E402
from tomoencoders_b200 import _device, _lib  # noqa: E402

L = _lib.lib()
n = 512
vol = blobs(n, 0.3, 1)
labels = torch.empty((n, n, n), dtype=torch.int32, device="cuda")
n_dev = torch.zeros(1, dtype=torch.int64, device="cuda")
ws = torch.empty(int(L.te_label_workspace_bytes(n ** 3)), dtype=torch.uint8, device="cuda")
for conn in (26, 6):
    _lib.check(L.te_label(vol.data_ptr(), _lib.TE_U8, _lib.i64x3((n, n, n)), conn, labels.data_ptr(), n_dev.data_ptr(), ws.data_ptr(),
                          _device.stream_ptr()), "te_label")
    torch.cuda.synchronize()
