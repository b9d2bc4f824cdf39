import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import _lib, synthetic
from oracle import planner_oracle as po

class Cfg:
    num_agents, nGraphFilterTaps, device = 10, 3, torch.device("cuda")

sd = po.init_state_dict(3, seed=1); po.randomize_bn_stats(sd)
m = gp.DecentralPlannerNet(Cfg()); m.load_state_dict(sd); m = m.cuda().eval(); m.set_feature_mode("tc")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, S = synthetic.make_batch(64, 10, 20, seed=3)
xt = torch.from_numpy(x).repeat(B // 64, 1, 1, 1, 1).cuda(); St = torch.from_numpy(S).repeat(B // 64, 1, 1).cuda()
_lib.set_debug_option("tc_timing", 1); lib = _lib.load(); out = (C.c_ulonglong * 32)()
with torch.no_grad():
    for _ in range(3):
        m.addGSO(St); m(xt)
    lib.gpp_debug_feature_tc_timing(out)
    m.addGSO(St); m(xt)
    lib.gpp_debug_feature_tc_timing(out)
v = list(out); t = max(v[18], 1)
print("tiles", t)
for L in range(6):
    print("layer %d: stage %7.0f  wait-mma %7.0f  epilogue %7.0f cycles/tile" % (L, v[3*L]/t, v[3*L+1]/t, v[3*L+2]/t))
print("total per tile %.0f" % (sum(v[:18]) / t))
it = max(v[24], 1)
print("producer per item: wait-doneA %.0f  gather/split/store %.0f  fence.proxy.async %.0f  arrive %.0f  (items %d)" % (v[20]/it, v[21]/it, v[22]/it, v[23]/it, it))
print("mma thread per item: wait-full %.0f  issue+commit %.0f" % (v[26]/it, v[27]/it))
