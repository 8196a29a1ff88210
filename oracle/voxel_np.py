"""numpy restatement of the voxel pre-processing on the hot path.  TEST INFRASTRUCTURE.

Follows /root/reference/tomo_encoders/misc/voxel_processing.py:81-104 and the clip in
neural_nets/surface_segmenter.py:272-275.
"""
import numpy as np


def rescale_data(data, min_val, max_val):
    """voxel_processing.py:81-89: (data - min) / (max - min + 1e-12), numpy dtype promotion rules."""
    eps = 1e-12
    return (data - min_val) / (max_val - min_val + eps)


def find_min_max(vol, sampling_factor):
    """voxel_processing.py:91-104: min / max over vol[::s, ::s, ::s]."""
    ss = slice(None, None, sampling_factor)
    sub = vol[ss, ss, ss]
    return sub.min(), sub.max()


def clip_rescale_f32(x, min_max, clip):
    """The arithmetic the product kernels implement: everything in float32 (SURVEY.md H6 -- the
    reference's own result depends on the numpy promotion of the input dtype; the comparison is made
    at the latent / mask level with the tolerances of BASELINE.json.north_star)."""
    x = np.asarray(x).astype(np.float32)
    if min_max is None:
        return x
    lo, hi = np.float32(min_max[0]), np.float32(min_max[1])
    if clip:
        x = np.clip(x, lo, hi)
    return ((x - lo) / (hi - lo + np.float32(1e-12))).astype(np.float32)
