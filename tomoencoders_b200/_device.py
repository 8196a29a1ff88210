"""Device-side plumbing: torch tensors are the handle type for HBM buffers and streams.

numpy in -> numpy out (host<->device copies inside the call, like the reference's TF path);
torch CUDA tensor in -> torch CUDA tensor out (like the reference's cupy-in / cupy-out behaviour,
structures/patches.py:759, structures/grid.py:324).
"""
import numpy as np
import torch

from . import _lib

_NP_TO_TE = {np.dtype(np.uint8): _lib.TE_U8, np.dtype(np.uint16): _lib.TE_U16,
             np.dtype(np.float16): _lib.TE_F16, np.dtype(np.float32): _lib.TE_F32}
_TORCH_TO_TE = {torch.uint8: _lib.TE_U8, torch.uint16: _lib.TE_U16, torch.float16: _lib.TE_F16,
                torch.float32: _lib.TE_F32, torch.bfloat16: _lib.TE_BF16}
_TE_TO_TORCH = {_lib.TE_U8: torch.uint8, _lib.TE_U16: torch.uint16, _lib.TE_F16: torch.float16,
                _lib.TE_F32: torch.float32, _lib.TE_BF16: torch.bfloat16}


def require_cuda():
    if not torch.cuda.is_available():
        raise _lib.TomoEncError("a CUDA device is required: tomoencoders_b200 has no CPU fallback")


def device():
    require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


def is_device_array(a):
    return isinstance(a, torch.Tensor) and a.is_cuda


def te_dtype(t):
    try:
        return _TORCH_TO_TE[t.dtype]
    except KeyError:
        raise TypeError("unsupported volume dtype %s (supported: uint8, uint16, float16, float32)" % (t.dtype,))


def torch_dtype(te_code):
    return _TE_TO_TORCH[te_code]


def to_device(a, non_blocking=True):
    """numpy array / torch CPU tensor / torch CUDA tensor -> contiguous torch CUDA tensor."""
    require_cuda()
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(device(), non_blocking=non_blocking)
        return t.contiguous()
    a = np.ascontiguousarray(a)
    if a.dtype not in _NP_TO_TE and a.dtype not in (np.dtype(np.int32), np.dtype(np.uint32), np.dtype(np.int64),
                                                    np.dtype(np.float64)):
        raise TypeError("unsupported array dtype %s" % (a.dtype,))
    return torch.from_numpy(a).to(device(), non_blocking=False)


def to_host_like(t, like):
    """Returns t in the array family of ``like`` (numpy for host input, torch CUDA for device input)."""
    if is_device_array(like):
        return t
    out = t.cpu()
    if isinstance(like, torch.Tensor):
        return out
    return out.numpy()


def to_supported(t):
    """Device tensor of any numeric dtype -> a dtype the kernels read (uint8 / uint16 / float16 / float32).

    The reference accepts whatever numpy holds: float64 is numpy's default (its own random_data_generator yields it,
    keras_processor.py:234-243) and Keras casts the network input to float32 anyway, so float64 / bfloat16 go to
    float32; signed and wide integers go to float32 as well (exact up to 2**24, far beyond any detector's range)."""
    if t.dtype in _TORCH_TO_TE and t.dtype != torch.bfloat16:
        return t
    if t.dtype == torch.bool:
        return t.to(torch.uint8)
    return t.to(torch.float32)


def stream_ptr():
    return torch.cuda.current_stream().cuda_stream


_SIDE_STREAMS = {}


def side_stream():
    """One extra stream per device for host<->device copies that overlap the kernels of the current stream."""
    d = torch.cuda.current_device()
    if d not in _SIDE_STREAMS:
        _SIDE_STREAMS[d] = torch.cuda.Stream(device=d)
    return _SIDE_STREAMS[d]


def as_host_tensor(a):
    """numpy array / torch CPU tensor -> torch CPU tensor sharing its memory (None if that is not possible)."""
    if isinstance(a, torch.Tensor):
        ok = not a.is_cuda and a.is_contiguous() and a.dtype in _TORCH_TO_TE and a.dtype != torch.bfloat16
        return a if ok else None
    if isinstance(a, np.ndarray) and a.flags.c_contiguous and a.dtype in _NP_TO_TE and a.flags.writeable:
        return torch.from_numpy(a)
    return None


_POINTS_CACHE = {}
_POINTS_CACHE_MAX_ROWS = 65536
_POINTS_CACHE_ENTRIES = 16


def points_to_device(points):
    """(n,3) uint32 coordinates -> int32-typed CUDA tensor holding the same bits.

    Small tables (Grid tilings are re-used call after call) are cached on the device, keyed by
    their contents, so that repeated extract / fill calls do not pay a synchronous upload."""
    p = np.ascontiguousarray(np.asarray(points), dtype=np.uint32)
    dev = device()
    if len(p) > _POINTS_CACHE_MAX_ROWS:
        return torch.from_numpy(p.view(np.int32)).to(dev)
    key = (dev.index, p.shape, p.tobytes())
    t = _POINTS_CACHE.get(key)
    if t is None:
        if len(_POINTS_CACHE) >= _POINTS_CACHE_ENTRIES:
            _POINTS_CACHE.pop(next(iter(_POINTS_CACHE)))
        t = torch.from_numpy(p.view(np.int32)).to(dev)
        _POINTS_CACHE[key] = t
    return t
