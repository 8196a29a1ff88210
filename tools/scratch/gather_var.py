import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import _device, _lib
from tools.gather_limits import timed_us
L, st = _lib.lib(), _device.stream_ptr()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush_r = torch.zeros(256 << 18, dtype=torch.int32, device="cuda")
def case(V, P, n, variants):
    vshape = (V,) * 3
    o8 = torch.empty((n, P, P, P), dtype=torch.uint8, device="cuda")
    v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device="cuda")
    vs_, p_ = _lib.i64x3(vshape), _lib.i32x3((P,) * 3)
    rg = np.random.default_rng(3)
    rp = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
    dp, dw = _device.points_to_device(rp), _device.points_to_device(np.full_like(rp, P))
    z, y = rp[:, 0].astype(np.int64), rp[:, 1].astype(np.int64)
    perm = torch.from_numpy(np.argsort(z * V + y, kind="stable").astype(np.int32)).cuda()
    for var in variants:
        os.environ["TOMOENC_GATHER_VARIANT"] = str(var)
        t, ta = timed_us(lambda: _lib.check(L.te_gather_perm(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                              _lib.TE_HINT_UNBINNED, perm.data_ptr(), st)), flush, flush_r)
        print("%d x %d^3 of %d^3, variant %s: %8.1f us (avg %8.1f)" % (n, P, V, var, t, ta), flush=True)
variants = [int(a) for a in sys.argv[1:]] or [0, 2, 3, 4]
case(512, 64, 16384, variants)
case(256, 32, 4096, variants)
