"""Seeded synthetic volumes for the BASELINE configs.  TEST INFRASTRUCTURE.

Blobs phantom: b = gaussian_filter(N(0,1), sigma), Y = b > quantile(b, 0.6),
X = Y + 0.1 N(0,1), min-max scaled.
"""
import numpy as np


def blobs(shape, seed_struct, seed_noise, sigma=4.0):
    from scipy.ndimage import gaussian_filter
    shape = tuple(shape)
    b = gaussian_filter(np.random.default_rng(seed_struct).standard_normal(shape, dtype=np.float32), sigma)
    y = b > np.quantile(b.ravel()[:: max(1, b.size // 2_000_000)], 0.6)
    x = y.astype(np.float32) + 0.1 * np.random.default_rng(seed_noise).standard_normal(shape, dtype=np.float32)
    x = (x - x.min()) / (x.max() - x.min())
    return x, y.astype(np.uint8)


def volume_u8(shape, seed_struct=1, seed_noise=2):
    x, y = blobs(shape, seed_struct, seed_noise)
    return np.round(x * 255.0).astype(np.uint8), y


def volume_f16(shape, seed_struct=4, seed_noise=5):
    x, y = blobs(shape, seed_struct, seed_noise)
    return x.astype(np.float16), y


def random_points(vol_shape, patch, n, seed):
    """semantics of structures/patches.py:379 (high exclusive): corner in [0, vol - patch)."""
    g = np.random.default_rng(seed)
    return np.stack([g.integers(0, vol_shape[a] - patch, n) for a in range(3)], axis=1).astype(np.uint32)


def cosine_volume(shape, dtype, seed):
    """Smooth seeded volume with a skewed histogram (separable cosines + noise); inputs of
    tests/golden/voxel_golden.npz."""
    g = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij")
    v = np.cos(z / 5.0) * np.cos(y / 7.0) + np.sin(x / 3.0) ** 2 + 0.3 * g.standard_normal(shape)
    v = (v - v.min()) / (v.max() - v.min())
    if np.dtype(dtype) == np.uint8:
        return (v * 255).astype(np.uint8)
    if np.dtype(dtype) == np.uint16:
        return (v * 40000 + 1000).astype(np.uint16)
    return (v * 3.0 - 1.0).astype(dtype)


def cosine_labels(shape, seed):
    g = np.random.default_rng(seed)
    z, y, x = np.meshgrid(*[np.arange(s, dtype=np.float64) for s in shape], indexing="ij")
    return (np.cos(z / 4.0) + np.cos(y / 5.0) + np.cos(x / 6.0) + 0.2 * g.standard_normal(shape)) > 0.3
