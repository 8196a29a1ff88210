"""te_pca_solve alone: time (CUDA events) and number of Jacobi sweeps (work[0]) for a few covariance spectra, d = 128."""
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from tomoencoders_b200 import _device, _lib  # noqa: E402

L = _lib.lib()
d, k, n = 128, 2, 4096
g = torch.Generator(device="cuda").manual_seed(0)
for name, scale in (("isotropic", torch.ones(d, device="cuda")), ("decaying 0.9^i", 0.9 ** torch.arange(d, device="cuda")),
                    ("rank 16", torch.cat([torch.ones(16, device="cuda"), torch.zeros(d - 16, device="cuda")]))):
    h = (torch.randn((n, d), device="cuda", generator=g) * scale) @ torch.linalg.qr(torch.randn((d, d), device="cuda", generator=g))[0]
    h = h.float().contiguous()
    mom = torch.zeros(1 + d + d * d, dtype=torch.float64, device="cuda")
    _lib.check(L.te_pca_accumulate(h.data_ptr(), n, d, mom.data_ptr(), _device.stream_ptr()), "acc")
    mean = torch.empty(d, dtype=torch.float64, device="cuda")
    comps = torch.empty((k, d), dtype=torch.float64, device="cuda")
    evar = torch.empty(k, dtype=torch.float64, device="cuda")
    work = torch.zeros(2 * d * d, dtype=torch.float64, device="cuda")
    def run():
        _lib.check(L.te_pca_solve(mom.data_ptr(), d, k, mean.data_ptr(), comps.data_ptr(), evar.data_ptr(), work.data_ptr(),
                                  _device.stream_ptr()), "solve")
    run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        run()
    b.record()
    torch.cuda.synchronize()
    cov = torch.cov(h.double().T)
    w, v = torch.linalg.eigh(cov)
    cosv = [abs(float((comps[c] @ v[:, -1 - c]))) for c in range(k)]
    print("%-16s %.3f ms per solve, %d sweeps, evar rel err %.1e, |cos| %s" %
          (name, a.elapsed_time(b) / 5, int(work[0].item()), float(((evar - w.flip(0)[:k]).abs() / w.flip(0)[:k]).max()), cosv))
