"""Processing orders x cache-policy variants of the row-group gather (cfg3-like case)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tomoencoders_b200 import _device, _lib
from tools.gather_limits import timed_us

def morton(z, y, x, bits=10):
    c = np.zeros(len(z), np.uint64)
    for b in range(bits):
        c |= ((x >> b) & 1).astype(np.uint64) << np.uint64(3 * b)
        c |= ((y >> b) & 1).astype(np.uint64) << np.uint64(3 * b + 1)
        c |= ((z >> b) & 1).astype(np.uint64) << np.uint64(3 * b + 2)
    return c

def main(vshape=(512,) * 3, P=64, n=16384, seed=22):
    L, st = _lib.lib(), _device.stream_ptr()
    v8 = torch.randint(0, 256, vshape, dtype=torch.uint8, device="cuda")
    o8 = torch.empty((n, P, P, P), dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    flush_r = torch.zeros(256 << 18, dtype=torch.int32, device="cuda")
    vs_, p_ = _lib.i64x3(vshape), _lib.i32x3((P,) * 3)
    rg = np.random.default_rng(seed)
    rp = np.stack([rg.integers(0, vshape[a] - P, n) for a in range(3)], axis=1).astype(np.uint32)
    dp, dw = _device.points_to_device(rp), _device.points_to_device(np.full_like(rp, P))
    z, y, x = (rp[:, i].astype(np.int64) for i in range(3))
    orders = {
        "index": np.arange(n),
        "(z,y)": np.argsort(z * vshape[1] + y, kind="stable"),
        "z only": np.argsort(z, kind="stable"),
        "morton": np.argsort(morton(z, y, x), kind="stable"),
        "morton/16": np.argsort(morton(z >> 4, y >> 4, x >> 4), kind="stable"),
        "(z/32,y/32,x)": np.lexsort((x, y >> 5, z >> 5)),
        "(z/64,y/64,x/64) blocks": np.lexsort((x >> 6, y >> 6, z >> 6)),
    }
    for var in ("0", "1", "2"):
        os.environ["TOMOENC_GATHER_VARIANT"] = var
        for name, o in orders.items():
            perm = torch.from_numpy(o.astype(np.int32)).cuda()
            t, ta = timed_us(lambda: _lib.check(L.te_gather_perm(v8.data_ptr(), 1, vs_, dp.data_ptr(), dw.data_ptr(), n, p_, o8.data_ptr(),
                                                                  _lib.TE_HINT_UNBINNED, perm.data_ptr(), st)), flush, flush_r)
            print("variant %s order %-26s %8.1f us (avg %8.1f)" % (var, name, t, ta), flush=True)

if __name__ == "__main__":
    main()
    print("cfg1")
    main((256,) * 3, 32, 4096, 21)
