import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import _device, _lib
L, st = _lib.lib(), _device.stream_ptr()
for V, P, n in ((128, 32, 64), (256, 64, 256)):
    vshape = (V,) * 3
    v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device="cuda")
    o8 = torch.zeros((n, P, P, P), dtype=torch.uint8, device="cuda")
    rg = np.random.default_rng(3)
    rp = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
    dp, dw = _device.points_to_device(rp), _device.points_to_device(np.full_like(rp, P))
    vs_, p_ = _lib.i64x3(vshape), _lib.i32x3((P,) * 3)
    rc = L.te_gather(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(), _lib.TE_HINT_UNBINNED, st)
    print("rc", rc, L.te_last_error() if rc else "")
    torch.cuda.synchronize()
    vh = v8.cpu().numpy(); oh = o8.cpu().numpy()
    bad = 0
    for i in range(n):
        z, y, x = (int(c) for c in rp[i])
        if not np.array_equal(oh[i], vh[z:z + P, y:y + P, x:x + P]): bad += 1
    print(V, P, n, "bad patches:", bad)
