"""GPU: the tcgen05 weight-gradient operator (te_tr_conv3_wgrad, csrc/wgrad_umma.cu) through the C ABI
against a plain PyTorch fp32 reference of the same op (autograd of conv3d).  The inputs are small
integers, exactly representable in fp16 / bf16, and every partial sum stays below 2^24, so the fp32
tensor-core accumulation is exact and the comparison is BIT-EXACT for any summation order.  Shapes
cover ragged tiles (H, W, D not multiples of the 16 x 8 x 16 work unit), channel-slice views
(coff / ctot), the 16-channel (SWIZZLE_32B) and 32-channel (SWIZZLE_64B) operand layouts, several
channel-chunk pairs per launch and accumulation into a non-zero dw."""
import ctypes

import numpy as np
import pytest
import torch

from tomoencoders_b200 import _device, _lib

pytestmark = pytest.mark.gpu



def _te_dtype(t):
    return _device.te_dtype(t)


def _view(t, coff, C):
    """te_tensor for channels [coff, coff + C) of an NDHWC tensor t (N, D, H, W, ctot)."""
    N, D, H, W, ctot = t.shape
    return _lib.TeTensor(t.data_ptr(), _te_dtype(t), ctot, coff, C, N, D, H, W)


def _reference(x, dy):
    """x (N,D,H,W,Ci), dy (N,D,H,W,Co) float32 on the GPU -> dw (3,3,3,Ci,Co) float32."""
    torch.backends.cudnn.allow_tf32 = False
    xx = x.permute(0, 4, 1, 2, 3).contiguous()
    w = torch.zeros((dy.shape[-1], x.shape[-1], 3, 3, 3), device=x.device, requires_grad=True)
    y = torch.nn.functional.conv3d(xx, w, padding=1)
    y.backward(dy.permute(0, 4, 1, 2, 3).contiguous())
    return w.grad.permute(2, 3, 4, 1, 0).contiguous()


CASES = [
    # N, D, H, W, x_ctot, x_coff, Ci, dy_ctot, dy_coff, Co
    (2, 16, 16, 8, 32, 0, 32, 32, 0, 32),
    (1, 20, 24, 24, 16, 0, 16, 64, 0, 64),
    (1, 8, 8, 8, 128, 0, 128, 64, 0, 64),
    (2, 16, 32, 16, 96, 0, 96, 32, 0, 32),
    (1, 12, 16, 40, 48, 16, 32, 96, 32, 64),
    (3, 4, 20, 12, 64, 0, 64, 32, 0, 32),
]


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("case", CASES)
def test_wgrad_operator_is_exact_on_integer_inputs(case, dtype):
    N, D, H, W, xct, xoff, Ci, dct, doff, Co = case
    dev = _device.device()
    g = torch.Generator(device=dev).manual_seed(N * 1000 + D * 100 + Ci)
    x = torch.randint(-3, 4, (N, D, H, W, xct), device=dev, generator=g).to(dtype)
    dy = torch.randint(-2, 3, (N, D, H, W, dct), device=dev, generator=g).to(dtype)
    dw_cin = Ci + 16                      # the slice lands at channel offset 16 of a wider dw
    dw0 = torch.randint(-5, 6, (27, dw_cin, Co), device=dev, generator=g).float()
    dw = dw0.clone()
    L = _lib.lib()
    tx, td = _view(x, xoff, Ci), _view(dy, doff, Co)
    _lib.check(L.te_tr_conv3_wgrad(ctypes.byref(tx), ctypes.byref(td), dw.data_ptr(), dw_cin, 16, None,
                                   _device.stream_ptr()), "te_tr_conv3_wgrad")
    torch.cuda.synchronize()
    ref = _reference(x[..., xoff:xoff + Ci].float(), dy[..., doff:doff + Co].float()).reshape(27, Ci, Co)
    got = dw - dw0
    assert torch.equal(got[:, :16], torch.zeros_like(got[:, :16]))        # channels outside the slice untouched
    assert torch.equal(got[:, 16:], ref), float((got[:, 16:] - ref).abs().max())


def test_wgrad_operator_random_values_close_to_fp32():
    """Real-valued bf16 inputs: fp32 accumulation in a different order -> relative error ~1e-6 of the norm."""
    dev = _device.device()
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.randn((2, 32, 32, 32, 32), device=dev, generator=g).to(torch.bfloat16)
    dy = torch.randn((2, 32, 32, 32, 64), device=dev, generator=g).to(torch.bfloat16)
    dw = torch.zeros((27, 32, 64), device=dev)
    tx, td = _view(x, 0, 32), _view(dy, 0, 64)
    _lib.check(_lib.lib().te_tr_conv3_wgrad(ctypes.byref(tx), ctypes.byref(td), dw.data_ptr(), 32, 0, None,
                                            _device.stream_ptr()), "te_tr_conv3_wgrad")
    ref = _reference(x.float(), dy.float()).reshape(27, 32, 64)
    assert float((dw - ref).norm() / ref.norm()) < 1e-5


def test_first_layer_wgrad_is_exact_on_integer_inputs():
    """Cin = 1 (register-tiled CUDA-core kernel): ragged 20 x 24 x 12 grid, Cout = 16, bf16 and fp16."""
    dev = _device.device()
    for dtype in (torch.bfloat16, torch.float16):
        g = torch.Generator(device=dev).manual_seed(11)
        x = torch.randint(-3, 4, (2, 20, 24, 12, 1), device=dev, generator=g).to(dtype)
        dy = torch.randint(-2, 3, (2, 20, 24, 12, 16), device=dev, generator=g).to(dtype)
        dw = torch.zeros((27, 1, 16), device=dev)
        tx, td = _view(x, 0, 1), _view(dy, 0, 16)
        _lib.check(_lib.lib().te_tr_conv3_wgrad(ctypes.byref(tx), ctypes.byref(td), dw.data_ptr(), 1, 0, None,
                                                _device.stream_ptr()), "te_tr_conv3_wgrad")
        ref = _reference(x.float(), dy.float()).reshape(27, 1, 16)
        assert torch.equal(dw, ref), float((dw - ref).abs().max())


@pytest.mark.parametrize("accumulate", [0, 1])
def test_maxpool_backward_vectorised_equals_scalar_kernel(accumulate):
    """The 8-channel kernel (16-bit tensors) against the scalar kernel (the same values as fp32 tensors):
    same first-argmax tie rule, so the gradients are identical, ties included (values on a coarse grid)."""
    dev = _device.device()
    g = torch.Generator(device=dev).manual_seed(5)
    x32 = torch.randint(-8, 9, (2, 8, 12, 16, 32), device=dev, generator=g).float() * 0.25
    gp32 = torch.randint(-8, 9, (2, 4, 6, 8, 32), device=dev, generator=g).float() * 0.5
    dx0 = torch.randint(-4, 5, x32.shape, device=dev, generator=g).float()
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        x, gp, dx = x32.to(dt), gp32.to(dt), dx0.to(dt).clone()
        tx, tg, tdx = _view(x, 0, 32), _view(gp, 0, 32), _view(dx, 0, 32)
        _lib.check(_lib.lib().te_tr_maxpool_bwd(ctypes.byref(tx), ctypes.byref(tg), ctypes.byref(tdx), 2, accumulate,
                                                _device.stream_ptr()), "te_tr_maxpool_bwd")
        outs.append(dx.float())
    assert torch.equal(outs[0], outs[1])
    # and against torch autograd where the window maximum is unique
    xx = x32.permute(0, 4, 1, 2, 3).contiguous().requires_grad_(True)
    torch.nn.functional.max_pool3d(xx, 2).backward(gp32.permute(0, 4, 1, 2, 3).contiguous())
    ref = xx.grad.permute(0, 2, 3, 4, 1) + (dx0 if accumulate else 0)
    w = x32.reshape(2, 4, 2, 6, 2, 8, 2, 32).permute(0, 1, 3, 5, 7, 2, 4, 6).reshape(2, 4, 6, 8, 32, 8)
    top2 = w.topk(2, dim=-1).values
    unique = (top2[..., 0] > top2[..., 1])
    unique = unique[:, :, None, :, None, :, None, :].expand(2, 4, 2, 6, 2, 8, 2, 32).reshape(2, 8, 12, 16, 32)
    assert torch.equal(outs[1][unique], ref[unique])


def test_fused_head_loss_matches_unfused_fp32_path():
    """te_tr_head_loss on bf16 tensors (one fused 128-bit-vectorised pass) against the fp32 tensors path."""
    dev = _device.device()
    g = torch.Generator(device=dev).manual_seed(9)
    shape = (2, 16, 16, 24, 32)
    a16 = torch.randn(shape, device=dev, generator=g).to(torch.bfloat16)
    w = torch.randn(32, device=dev, generator=g) * 0.3
    b = torch.tensor([0.1], device=dev)
    y = (torch.rand(shape[:-1], device=dev, generator=g) < 0.4).to(torch.uint8)
    nvox = y.numel()
    res = []
    for dt in (torch.float32, torch.bfloat16):
        a = a16.to(dt)
        da = torch.zeros_like(a)
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        dwb = torch.zeros(64, dtype=torch.float64, device=dev)
        dl = torch.zeros(nvox, device=dev)
        probs = torch.zeros(nvox, device=dev)
        ta, tda = _view(a, 0, 32), _view(da, 0, 32)
        _lib.check(_lib.lib().te_tr_head_loss(ctypes.byref(ta), w.data_ptr(), b.data_ptr(), y.data_ptr(), ctypes.byref(tda),
                                              loss.data_ptr(), dwb.data_ptr(), float(nvox), dl.data_ptr(), probs.data_ptr(),
                                              _device.stream_ptr()), "te_tr_head_loss")
        res.append((float(loss), dwb.clone(), dl.clone(), probs.clone(), da.float()))
    (l0, d0, dl0, p0, da0), (l1, d1, dl1, p1, da1) = res
    assert abs(l0 - l1) < 1e-6 * abs(l0)
    torch.testing.assert_close(d1[:33], d0[:33], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(dl1, dl0, rtol=1e-5, atol=1e-12)
    torch.testing.assert_close(p1, p0, rtol=1e-5, atol=1e-7)
    torch.testing.assert_close(da1, da0, rtol=2 ** -7, atol=1e-12)       # da is stored in bf16
