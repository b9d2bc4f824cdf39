import sys, ctypes as C, torch
sys.path.insert(0, "/root/repo")
from gnn_pathplanning_b200 import _lib
import torch.nn.functional as Fn
lib = _lib.load()
lib.gpp_debug_train_kernel.argtypes = [C.c_int] + [C.c_void_p]*4 + [C.c_int]*4 + [C.c_void_p]
st = torch.cuda.current_stream().cuda_stream
torch.manual_seed(0)
for (M, Cin, Cout, H) in [(80, 32, 64, 5), (640, 64, 128, 2), (40, 3, 32, 11), (160, 32, 32, 5)]:
    x = torch.randn(M, Cin, H, H, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.1; b = torch.randn(Cout, device="cuda")
    y = torch.empty(M, Cout, H, H, device="cuda")
    lib.gpp_debug_train_kernel(0, x.data_ptr(), w.data_ptr(), b.data_ptr(), y.data_ptr(), M, Cin, Cout, H, st)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        ref = Fn.conv2d(x, w, b, padding=1)
    e0 = ((y - ref).abs().max() / ref.abs().max()).item()
    dz = torch.randn(M, Cout, H, H, device="cuda"); dx = torch.empty(M, Cin, H, H, device="cuda")
    lib.gpp_debug_train_kernel(1, dz.data_ptr(), w.data_ptr(), None, dx.data_ptr(), M, Cin, Cout, H, st)
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        refdx = torch.ops.aten.convolution_backward(dz, x, w, [Cout], [1,1],[1,1],[1,1], False, [0,0], 1, [True, False, False])[0]
    e1 = ((dx - refdx).abs().max() / refdx.abs().max()).item()
    a = torch.relu(torch.randn(M, Cout, H, H, device="cuda")).requires_grad_(True)
    p = Fn.max_pool2d(a, 2); dp = torch.randn_like(p); p.backward(dp)
    da = torch.empty_like(a)
    lib.gpp_debug_train_kernel(2, a.data_ptr(), dp.contiguous().data_ptr(), None, da.data_ptr(), M, Cin, Cout, H, st)
    e2 = (da - a.grad).abs().max().item()
    print("M=%d Cin=%d Cout=%d H=%d: conv_fwd %.1e  conv_bwd_input %.1e  pool_bwd %.1e" % (M, Cin, Cout, H, e0, e1, e2))
