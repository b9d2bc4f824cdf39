"""clock64 phase timers of feature_mma_kernel (block 0). usage: python profiles/fm_phase_timing.py [episodes]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import _lib, synthetic
from oracle import planner_oracle as po


class Cfg:
    num_agents, nGraphFilterTaps, device = 10, 3, torch.device("cuda")


B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sd = po.init_state_dict(3, seed=1); po.randomize_bn_stats(sd)
m = gp.DecentralPlannerNet(Cfg()); m.load_state_dict(sd); m = m.cuda().eval(); m.set_feature_mode("mma")
x, S = synthetic.make_batch(64, 10, 20, seed=3)
reps = (B + 63) // 64
xt = torch.from_numpy(x).repeat(reps, 1, 1, 1, 1)[:B].cuda(); St = torch.from_numpy(S).repeat(reps, 1, 1)[:B].cuda()
lib = _lib.load()
with torch.no_grad():
    for _ in range(3):
        m.addGSO(St); m(xt)
    torch.cuda.synchronize()
    _lib.set_debug_option("tc_timing", 1)
    for _ in range(5):
        m.addGSO(St); m(xt)
    torch.cuda.synchronize()
out = (C.c_ulonglong * 40)()
_lib.check(lib.gpp_debug_feature_mma_timing(out))
_lib.set_debug_option("tc_timing", 0)
t = [int(v) for v in out]
tiles = max(1, t[18])
names = ["conv0", "conv1", "conv2", "conv3", "conv4", "mlp"]
print("episodes %d (agents %d), tiles of block 0 over 5 launches: %d; cycles per tile" % (B, B * 10, tiles))
print("%-6s %12s %12s %12s | %12s %12s" % ("layer", "mma:wait-in", "mma:wait-w", "mma:issue", "epi:wait-acc", "epi:work"))
for L in range(6):
    print("%-6s %12.0f %12.0f %12.0f | %12.0f %12.0f" % (names[L], t[L] / tiles, t[6 + L] / tiles, t[12 + L] / tiles,
                                                        t[20 + L] / tiles, t[26 + L] / tiles))
print("conv0: mma waiting for a free pair buffer %.0f, epilogue TMEM load part %.0f" % (t[32] / tiles, t[33] / tiles))
print("convert %.0f   mma total %.0f   epilogue total %.0f" % (t[19] / tiles, sum(t[0:18]) / tiles, sum(t[19:32]) / tiles))
