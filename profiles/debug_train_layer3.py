import sys, torch, numpy as np
sys.path.insert(0, "/root/repo")
import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import synthetic, planner as pl
from oracle import planner_oracle as po
class Cfg:
    def __init__(s, n, k): s.num_agents, s.nGraphFilterTaps, s.device = n, k, torch.device("cuda")
B, N, K = 64, 10, 3
sd = po.init_state_dict(K, seed=11); po.randomize_bn_stats(sd, seed=3)
x, S = synthetic.make_batch(B, N, 20, seed=21)
tgt = torch.from_numpy(synthetic.random_targets(B, N, seed=5)).cuda()
xt, St = torch.from_numpy(x).cuda(), torch.from_numpy(S).cuda()
res = []
acts = {}
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
a, b = res[0]["ConvLayers.12.bias"], res[1]["ConvLayers.12.bias"]
d = np.abs(a - b); print("bn3.bias: max|ref| %.3e; top diffs idx" % np.abs(b).max(), np.argsort(-d)[:6], d[np.argsort(-d)[:6]], "ref vals", b[np.argsort(-d)[:6]], "native", a[np.argsort(-d)[:6]])
a, b = res[0]["ConvLayers.11.weight"], res[1]["ConvLayers.11.weight"]
d = np.abs(a - b).reshape(64, -1).max(1); print("conv3.w per-out-channel max diff: top", np.argsort(-d)[:6], d[np.argsort(-d)[:6]], "max|ref| %.3e" % np.abs(b).max())
a, b = res[0]["ConvLayers.12.weight"], res[1]["ConvLayers.12.weight"]
d = np.abs(a - b); print("bn3.weight: top", np.argsort(-d)[:4], d[np.argsort(-d)[:4]], "max|ref| %.3e" % np.abs(b).max())
# statistics of layer-3 BN groups: variance per (agent, channel) from the torch path activations
m = gp.DecentralPlannerNet(Cfg(N, K)); m.load_state_dict(sd); m = m.cuda().train()
h = xt.reshape(B * N, 3, 11, 11)
with torch.no_grad():
    for l, ci in enumerate(pl._CONV_IDX):
        conv, bn = m.ConvLayers[ci], m.ConvLayers[ci + 1]
        z = pl._Conv3x3Fp32.apply(h, conv.weight, conv.bias)
        if l == 3:
            v = z.view(B, N, 64, 4).var(dim=(0, 3), unbiased=False)   # [N,64]
            print("layer3 group variance: min %.3e, #<1e-6: %d, #==0: %d; channels with tiny var:" % (v.min().item(), int((v < 1e-6).sum()), int((v == 0).sum())), torch.nonzero((v < 1e-6).any(0)).flatten().tolist())
        h = torch.relu(m._bn_per_agent(z, bn, N))
        if l % 2 == 0: h = torch.nn.functional.max_pool2d(h, 2)
