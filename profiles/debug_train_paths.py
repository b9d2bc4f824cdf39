import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import synthetic, planner as pl
from oracle import planner_oracle as po

class Cfg:
    def __init__(s, n, k): s.num_agents, s.nGraphFilterTaps, s.device = n, k, torch.device("cuda")

def run(B, N, K=3):
    sd = po.init_state_dict(K, seed=11); po.randomize_bn_stats(sd, seed=3)
    x, S = synthetic.make_batch(B, N, 20, seed=21)
    tgt = torch.from_numpy(synthetic.random_targets(B, N, seed=5)).cuda()
    xt, St = torch.from_numpy(x).cuda(), torch.from_numpy(S).cuda()
    res = []
    for native in (True, False):
        m = gp.DecentralPlannerNet(Cfg(N, K)); m.load_state_dict(sd); m = m.cuda().train()
        m.addGSO(St)
        if native:
            out = m(xt)
        else:
            m.GFL[0].addGSO(m.S)
            out = list(m._forward_autograd(xt, m.S).unbind(0))
        loss = po.planner_loss(out, tgt); loss.backward()
        res.append({n: p.grad.detach().cpu().numpy() for n, p in m.named_parameters()})
    worst = {}
    for n in res[0]:
        a, b = res[0][n], res[1][n]
        worst[n] = float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    bad = {k: "%.1e" % v for k, v in worst.items() if v > 1e-4 and not (k.endswith("bias") and k.split(".")[1] in ("0","4","7","11","14"))}
    print("B=%d N=%d: bad:" % (B, N), bad)

for B, N in [(4, 10), (8, 10), (16, 10), (64, 10), (64, 1), (64, 2), (16, 3)]:
    run(B, N)
