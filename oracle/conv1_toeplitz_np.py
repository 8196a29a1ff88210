"""Numpy model of the Toeplitz formulation of the first layer (tomoencoders_b200/csrc/conv1_tz.cu).  TEST INFRASTRUCTURE.

Restates, index for index, what the kernel's operands contain -- aligned 8-wide x "planes" with one halo plane on either
side, the two x phases of a 4-wide output tile, and the banded weight tiles B[phase][dz, dy][n = (xo, co)][k] -- and
evaluates the nine K = 16 products per tile in float64.  tests/test_oracle_nets_np.py checks it against the direct
convolution of oracle/nets_np.py (Conv3D(f, 3, 'same') on one channel, /root/reference/tomo_encoders/neural_nets/
Unet3D.py:117-122), which pins the decomposition itself (the index algebra of the kernel) on the CPU."""
import numpy as np

XT = 4      # output x positions per tile


def banded_weights(w):
    """w (3, 3, 3, 1, Co) Keras kernel -> B (2 phases, 9 (dz, dy), 16 K slots, XT * Co columns).
    Phase 0 (x0 = 8 g): K slot k holds input x0 - 8 + k; phase 1 (x0 = 8 g + 4): input x0 - 4 + k."""
    co = w.shape[-1]
    B = np.zeros((2, 9, 16, XT * co))
    for ph in range(2):
        for tap in range(9):
            for xo in range(XT):
                for dx in range(3):
                    k = (4 if ph else 8) + xo - 1 + dx                   # slot of input offset xo + dx - 1 relative to x0
                    B[ph, tap, k, xo * co:(xo + 1) * co] = w[tap // 3, tap % 3, dx, 0, :]
    return B


def conv1_toeplitz(x, w, b):
    """x (N, D, H, W) with W a multiple of 8, w (3, 3, 3, 1, Co), b (Co,) -> (N, D, H, W, Co) float64."""
    N, D, H, W = x.shape
    assert W % 8 == 0
    co = w.shape[-1]
    G = W // 8
    # planes[p] = x[..., 8 (p - 1) : 8 p] zero-padded in z, y and at the two out-of-range planes (what the TMA boxes deliver)
    xp = np.zeros((N, D + 2, H + 2, (G + 2) * 8))
    xp[:, 1:D + 1, 1:H + 1, 8:8 + W] = x
    B = banded_weights(np.asarray(w, dtype=np.float64))
    out = np.zeros((N, D, H, W, co))
    for t in range(W // XT):
        ph, g = t & 1, t >> 1
        p0 = g + ph                                                      # planes (p0, p0 + 1) = the K = 16 window of this tile
        acc = np.zeros((N, D, H, XT * co))
        for tap in range(9):
            dz, dy = tap // 3, tap % 3
            A = xp[:, dz:dz + D, dy:dy + H, 8 * p0:8 * p0 + 16]          # rows (z, y) shifted by the tap, 16 input x positions
            acc += A @ B[ph, tap]
        out[:, :, :, XT * t:XT * (t + 1), :] = acc.reshape(N, D, H, XT, co)
    return out + b
