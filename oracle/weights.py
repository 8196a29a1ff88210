"""Seeded weights in Keras ``get_weights()`` order and layout.  TEST INFRASTRUCTURE.

TensorFlow cannot be installed here, so "random-init weights" are drawn by this module (seeded
numpy) and loaded, unchanged, by both the oracle networks (oracle/nets_torch.py) and the product
(``processor.set_weights``).  Layouts:

  Conv3D            kernel (kz,ky,kx,Cin,Cout), bias (Cout)
  Conv3DTranspose   kernel (kz,ky,kx,Cout,Cin), bias (Cout)
  Dense             kernel (in,out), bias (out)
  BatchNormalization gamma, beta, moving_mean, moving_variance  (each (C,))

Each spec entry is ``(kind, shape)`` with kind in {"conv_k","conv_b","convT_k","convT_b","dense_k",
"dense_b","bn_gamma","bn_beta","bn_mean","bn_var"}.  Kernels are He-uniform for LeakyReLU(0.2)
(``U(+-sqrt(3 g^2 / fan_in))``, ``g^2 = 2/(1+0.2^2)``, ``fan_in = k^3 Cin``) instead of Keras' default
glorot-uniform, so that activations stay O(1) through 17 layers and bf16/fp16 error is representative
(SURVEY.md appendix A.2 leaves the scale to the oracle); biases are small non-zero normals and BN
statistics are randomised so that bias handling and BN folding are exercised.
"""
import numpy as np

LRELU_GAIN = float(np.sqrt(2.0 / (1.0 + 0.2 ** 2)))


def draw(spec, seed, gain=LRELU_GAIN):
    rng = np.random.default_rng(seed)
    out = []
    for kind, shape in spec:
        shape = tuple(int(s) for s in shape)
        if kind in ("conv_k", "convT_k"):
            k3 = int(np.prod(shape[:3]))
            if kind == "conv_k":
                fan_in = k3 * shape[3]
            else:
                # each output voxel of a k<=s transposed conv sees exactly one tap: fan_in is Cin
                fan_in = shape[4]
            lim = gain * np.sqrt(3.0 / fan_in)
            out.append(rng.uniform(-lim, lim, shape).astype(np.float32))
        elif kind == "dense_k":
            lim = gain * np.sqrt(3.0 / shape[0])
            out.append(rng.uniform(-lim, lim, shape).astype(np.float32))
        elif kind in ("conv_b", "convT_b", "dense_b"):
            out.append(rng.normal(0.0, 0.05, shape).astype(np.float32))
        elif kind == "bn_gamma":
            out.append(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        elif kind == "bn_beta":
            out.append(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif kind == "bn_mean":
            out.append(rng.normal(0.0, 0.1, shape).astype(np.float32))
        elif kind == "bn_var":
            out.append(rng.uniform(0.5, 1.5, shape).astype(np.float32))
        else:
            raise ValueError(kind)
    return out
