import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import _device, _lib
from tools.gather_limits import timed_us
L, st = _lib.lib(), _device.stream_ptr()
n, P = 16384, 64
o8 = torch.empty((n, P, P, P), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush_r = torch.zeros(256 << 18, dtype=torch.int32, device="cuda")
V = 128
vshape = (V,) * 3
v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device="cuda")
vs_, p_ = _lib.i64x3(vshape), _lib.i32x3((P,) * 3)
rg = np.random.default_rng(3)
rnd = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
same = np.full_like(rnd, 0); same[:] = (3, 5, 7)
same0 = np.zeros_like(rnd)
ident = np.arange(n, dtype=np.int32)
shuf = rg.permutation(n).astype(np.int32)
for cname, rp in (("same (3,5,7)", same), ("same (0,0,0)", same0), ("random", rnd), ("random z,y; x=0", np.concatenate([rnd[:, :2], np.zeros((n, 1), np.uint32)], 1)),
                  ("z=y=0, x random", np.concatenate([np.zeros((n, 2), np.uint32), rnd[:, 2:]], 1))):
    for pname, perm in (("identity", ident), ("shuffled", shuf)):
        dp, dw = _device.points_to_device(np.ascontiguousarray(rp)), _device.points_to_device(np.full_like(rp, P))
        dperm = torch.from_numpy(perm).cuda()
        t, ta = timed_us(lambda: _lib.check(L.te_gather_perm(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                              _lib.TE_HINT_UNBINNED, dperm.data_ptr(), st)), flush, flush_r)
        print("corners %-18s perm %-9s %8.1f us (avg %8.1f)" % (cname, pname, t, ta), flush=True)
