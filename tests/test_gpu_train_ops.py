"""Single kernels of the native training path (gpp_debug_train_kernel) against float64 torch references:
the row-tiled convolution kernels of the planner's map sizes (11, 5, 2) and the generic-size fallbacks."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as Fn

pytestmark = pytest.mark.gpu
TOL = 1e-5


def _rel(a, b):
    return ((a.double() - b).abs().max() / b.abs().max()).item()


@pytest.mark.parametrize("M,Cin,Cout,H", [(40, 3, 32, 11), (160, 32, 32, 5), (80, 32, 64, 5), (640, 64, 64, 2),
                                          (640, 64, 128, 2), (7, 64, 64, 2), (1, 3, 5, 2), (7, 5, 9, 4), (3, 2, 3, 7)])
def test_conv_and_pool_kernels(M, Cin, Cout, H):
    from gnn_pathplanning_b200 import _lib
    lib = _lib.load()
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cuda").manual_seed(M * 131 + H)
    x = torch.randn(M, Cin, H, H, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) * 0.1
    b = torch.randn(Cout, device="cuda", generator=g)
    # forward
    y = torch.empty(M, Cout, H, H, device="cuda")
    assert lib.gpp_debug_train_kernel(0, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, Cin, Cout, H, st) == 0
    ref = Fn.conv2d(x.double(), w.double(), b.double(), padding=1)
    assert _rel(y, ref) <= TOL
    # input gradient: the transposed convolution of dz with the flipped filters
    dz = torch.randn(M, Cout, H, H, device="cuda", generator=g)
    dx = torch.empty(M, Cin, H, H, device="cuda")
    assert lib.gpp_debug_train_kernel(1, dz.data_ptr(), w.data_ptr(), None, dx.data_ptr(), M, Cin, Cout, H, st) == 0
    refdx = Fn.conv_transpose2d(dz.double(), w.double(), padding=1)
    assert _rel(dx, refdx) <= TOL
    # max-pool gradient (exact: routing only)
    if H >= 2:
        a = torch.relu(torch.randn(M, Cout, H, H, device="cuda", generator=g)).requires_grad_(True)
        p = Fn.max_pool2d(a, 2)
        dp = torch.randn(p.shape, device="cuda", generator=g)
        p.backward(dp)
        da = torch.empty_like(a)
        assert lib.gpp_debug_train_kernel(2, a.data_ptr(), dp.contiguous().data_ptr(), None, da.data_ptr(), M, Cin, Cout,
                                          H, st) == 0
        assert torch.equal(da, a.grad)       # ties (zeros after the ReLU) go to the first maximum, as in torch
    # weight + bias gradient (per-chunk partial sums + fixed-order chunk reduction, as the training step runs it)
    out = torch.empty(Cout * Cin * 9 + Cout, device="cuda")
    assert lib.gpp_debug_train_kernel(3, dz.data_ptr(), x.data_ptr(), None, out.data_ptr(), M, Cin, Cout, H, st) == 0
    xd = x.double().requires_grad_(False)
    wd = w.double().clone().requires_grad_(True)
    bd = b.double().clone().requires_grad_(True)
    (Fn.conv2d(xd, wd, bd, padding=1) * dz.double()).sum().backward()
    assert _rel(out[:Cout * Cin * 9].view(Cout, Cin, 3, 3), wd.grad) <= TOL
    assert _rel(out[Cout * Cin * 9:], bd.grad) <= TOL
