"""Voxel pre-processing helpers on the hot path (drop-in for the two functions of
/root/reference/tomo_encoders/misc/voxel_processing.py the path uses).

``_find_min_max`` runs on the GPU (csrc/gather_scatter.cu); ``_rescale_data`` keeps the reference's
host arithmetic for callers that rescale arrays themselves -- inside predict_patches /
predict_embeddings the rescale is fused into the gather kernel instead (te_gather_norm).
"""
import torch

from .. import _device, _lib


def _rescale_data(data, min_val, max_val):
    """voxel_processing.py:81-89: (data - min) / (max - min + 1e-12) in the caller's array family."""
    eps = 1e-12
    return (data - min_val) / (max_val - min_val + eps)


def _find_min_max(vol, sampling_factor, TIMEIT=False):
    """voxel_processing.py:91-104: (min, max) of vol[::s, ::s, ::s]; returns Python floats."""
    t = _device.to_device(vol)
    if t.dim() != 3:
        raise ValueError("expected a 3-D volume")
    out = torch.empty(2, dtype=torch.float32, device=t.device)
    scratch = torch.empty(2048, dtype=torch.float32, device=t.device)
    L = _lib.lib()
    _lib.check(L.te_minmax_strided(t.data_ptr(), _device.te_dtype(t), _lib.i64x3(t.shape), int(sampling_factor),
                                   out.data_ptr(), scratch.data_ptr(), _device.stream_ptr()), "te_minmax_strided")
    lo, hi = out.tolist()
    return lo, hi
