"""Generates tests/golden/voxel_golden.npz by running the REFERENCE's own misc/voxel_processing.py
functions (imported unmodified through oracle/ref_loader.py; needs /root/reference, so it only runs in
the build container) on seeded inputs.  normalize_volume_gpu needs cupy staging buffers: a numpy-backed
shim provides cp.zeros / .set / .get / cp.cuda.Stream (with cupy's dtype / shape check in .set).

    python tests/golden/make_voxel_golden.py
"""
import os
import sys
import types

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from oracle import ref_loader  # noqa: E402
from oracle.phantom import cosine_labels as labels, cosine_volume as blobs  # noqa: E402

DTYPES = (("u8", np.uint8), ("u16", np.uint16), ("f16", np.float16), ("f32", np.float32))


class _Arr(np.ndarray):
    def set(self, a):
        if a.dtype != self.dtype or a.shape != self.shape:
            raise ValueError("Shape or dtype mismatch")     # cupy.ndarray.set
        self[...] = a

    def get(self):
        return np.asarray(self)


class _Stream:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def synchronize(self):
        pass


def main():
    ref = ref_loader.load()
    V = ref.ref_voxel
    V.cp.zeros = lambda shape, dtype=np.float32: np.zeros(shape, dtype=dtype).view(_Arr)
    V.cp.float32 = np.float32
    V.cp.cuda = types.SimpleNamespace(Stream=_Stream)
    out = {}
    shape = (24, 40, 40)
    for name, dt in DTYPES:
        vol = blobs(shape, dt, 200)
        for f in (1, 2, 3):
            out["ac_%s_f%d" % (name, f)] = np.array(V.modified_autocontrast(vol, s=0.01, normalize_sampling_factor=f), dtype=np.float64)
        out["ac_%s_tuple" % name] = np.array(V.modified_autocontrast(vol, s=(0.05, 0.1), normalize_sampling_factor=2), dtype=np.float64)
        out["ac_%s_1d" % name] = np.array(V.modified_autocontrast(vol.reshape(-1)[:5001], s=0.02), dtype=np.float64)
        v4 = np.stack([blobs((16, 16, 16), dt, 201 + i) for i in range(3)])
        out["ac_%s_4d" % name] = np.array(V.modified_autocontrast(v4, s=0.05, normalize_sampling_factor=2), dtype=np.float64)
        # np.histogram is the third-party arithmetic under modified_autocontrast: pin its counts too
        out["hist_%s" % name] = np.histogram(vol[::2, ::2, ::2].reshape(-1), bins=500)[0]
        masked = vol.copy()
        V.cylindrical_mask(masked, 0.8, mask_val=3)
        out["cyl_%s" % name] = masked
        out["cylvals_%s" % name] = V.get_values_cyl_mask(vol, 0.9)
        out["edge_%s" % name] = V.edge_map((vol > vol.mean()).astype(dt))
    out["edge_bool"] = V.edge_map(labels(shape, 210).astype(np.uint8))
    m1 = blobs(shape, np.float32, 200)
    V.cylindrical_mask(m1, 1.0, mask_val=np.min(m1[::2, ::2, ::2]))
    out["cyl_f32_full"] = m1
    out["cylvals_f32_sub"] = V.get_values_cyl_mask(blobs(shape, np.float32, 200)[::2, ::2, ::2], 1.0)
    for f in (1, 4):
        nv = blobs(shape, np.float32, 220)
        out["norm_f32_f%d" % f] = V.normalize_volume_gpu(nv, chunk_size=8, normalize_sampling_factor=f)
    path = os.path.join(os.path.dirname(__file__), "voxel_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
