"""GPU: the first layer of every graph (Conv3D(f, 3, 'same') on ONE input channel, Unet3D.py:117-122,
autoencoders.py:168-173) through the C ABI (te_tr_conv3) against a plain PyTorch fp32 conv3d of the same op.
The inputs, weights and biases are small integers: every product and partial sum is exactly representable,
so the tensor-core kernels (Toeplitz formulation csrc/conv1_tz.cu, im2col formulation csrc/conv1_umma.cu) and
the CUDA-core kernel must all be BIT-EXACT for any summation order.  Shapes cover both x phases of the
Toeplitz tiles, several work units per CTA (more units than SMs), patches that are one unit thin, and
shapes the Toeplitz kernel refuses (they must fall back, not fail)."""
import ctypes

import pytest
import torch

from tomoencoders_b200 import _device, _lib

pytestmark = pytest.mark.gpu


def _desc(t):
    N, D, H, W, C = t.shape
    return _lib.TeTensor(t.data_ptr(), _device.te_dtype(t), C, 0, C, N, D, H, W)


# N, D, H, W, Cout, expected kernel kind (2 Toeplitz, 1 im2col UMMA, 0 CUDA cores)
CASES = [
    (2, 16, 8, 8, 16, 2),        # one unit per patch, two x tiles (one per phase)
    (3, 32, 32, 32, 16, 2),      # BASELINE configs[0] patch size
    (5, 64, 64, 64, 16, 2),      # configs[1] patch size, 1280 units > 148 CTAs
    (1, 16, 24, 40, 16, 2),      # non-cubic
    (2, 16, 16, 72, 16, 1),      # W > 64: im2col kernel
    (2, 20, 12, 12, 16, 1),      # extents that are no multiples of the unit: im2col kernel
    (2, 16, 16, 16, 32, 1),      # Cout = 32: im2col kernel
    (1, 8, 8, 8, 48, 0),         # CUDA cores
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("case", CASES)
def test_first_layer_is_exact_on_integer_inputs(case, dtype):
    N, D, H, W, Co, kind = case
    dev = _device.device()
    L = _lib.lib()
    g = torch.Generator(device=dev).manual_seed(D * 1000 + W * 10 + Co)
    x = torch.randint(-3, 4, (N, D, H, W, 1), device=dev, generator=g).to(dtype)
    w = torch.randint(-2, 3, (27, 1, Co), device=dev, generator=g).float()
    b = torch.randint(-4, 5, (Co,), device=dev, generator=g).float()
    y = torch.full((N, D, H, W, Co), 7.0, device=dev, dtype=dtype)
    xd, yd = _desc(x), _desc(y)
    assert L.te_conv1_kernel_kind(ctypes.byref(xd), ctypes.byref(yd)) == kind
    _lib.check(L.te_tr_conv3(ctypes.byref(xd), None, w.data_ptr(), b.data_ptr(), ctypes.byref(yd), None,
                             _device.stream_ptr()), "te_tr_conv3")
    torch.cuda.synchronize()
    torch.backends.cudnn.allow_tf32 = False
    wk = w.reshape(3, 3, 3, 1, Co).permute(4, 3, 0, 1, 2).contiguous()
    ref = torch.nn.functional.conv3d(x.float().permute(0, 4, 1, 2, 3), wk, b, padding=1).permute(0, 2, 3, 4, 1)
    assert torch.equal(y.float(), ref), "max |diff| %g" % (y.float() - ref).abs().max().item()


def test_first_layer_of_the_benchmark_unet_runs_the_toeplitz_kernel():
    from tomoencoders_b200.neural_nets import SurfaceSegmenter
    fe = SurfaceSegmenter(model_initialization="define-new", descriptor_tag="t", mode="fp16", n_filters=[16, 32, 64],
                          n_blocks=3, activation="lrelu", batch_norm=True, isconcat=[True] * 3, pool_size=[2, 2, 2])
    net = fe.models["segmenter"].net_for((64,) * 3, 4)
    names = [row[1] for row in net.layer_times()]
    assert names[0] == "conv1_tz_kernel<16>", names[0]
