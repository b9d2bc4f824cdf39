import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import synthetic, planner as pl
from oracle import planner_oracle as po
import torch.nn.functional as Fn

class Cfg:
    def __init__(s, n, k): s.num_agents, s.nGraphFilterTaps, s.device = n, k, torch.device("cuda")

def run(B, N, K=3):
    sd = po.init_state_dict(K, seed=11); po.randomize_bn_stats(sd, seed=3)
    x, S = synthetic.make_batch(B, N, 20, seed=21)
    xt = torch.from_numpy(x).cuda()
    m = gp.DecentralPlannerNet(Cfg(N, K)); m.load_state_dict(sd); m = m.cuda().train()
    h = xt.reshape(B * N, 3, 11, 11)
    msg = []
    with torch.no_grad():
        for l, ci in enumerate(pl._CONV_IDX):
            conv, bn = m.ConvLayers[ci], m.ConvLayers[ci + 1]
            h = pl._Conv3x3Fp32.apply(h, conv.weight, conv.bias)
            h = torch.relu(m._bn_per_agent(h, bn, N))
            if l % 2 == 0:
                H = h.shape[-1]; Hp = H // 2
                win = h[:, :, :2*Hp, :2*Hp].reshape(h.shape[0], h.shape[1], Hp, 2, Hp, 2).permute(0,1,2,4,3,5).reshape(h.shape[0], h.shape[1], Hp, Hp, 4)
                top2 = win.topk(2, dim=-1).values
                ties = ((top2[..., 0] == top2[..., 1]) & (top2[..., 0] > 0))
                # which index does torch pick vs first-max
                _, idx = Fn.max_pool2d(h, 2, return_indices=True)
                first = win.argmax(-1)   # torch.argmax returns first max? not guaranteed; compute manually
                mx = win.max(-1, keepdim=True).values
                firstmax = (win == mx).float().argmax(-1)  # first True
                ys = (idx // H) % 2; xs = (idx % H) % 2
                picked = ys * 2 + xs
                diff = (picked != firstmax) & ties
                msg.append("L%d ties=%d pick!=first=%d" % (l, int(ties.sum()), int(diff.sum())))
                h = Fn.max_pool2d(h, 2)
    print("B=%d N=%d:" % (B, N), "; ".join(msg))

for B, N in [(4, 10), (8, 10), (16, 10), (64, 10), (64, 2)]:
    run(B, N)
