"""One training step (forward + fused loss + backward + Adam) at BASELINE configs 3 and 5 (per-GPU shard), device-resident,
CUDA events; CPU oracle step timed beside it (bounded).
usage: python profiles/train_microbench.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import synthetic
from oracle import planner_oracle as po


class Cfg:
    def __init__(self, n, k):
        self.num_agents, self.nGraphFilterTaps, self.device = n, k, torch.device("cuda")


for (N, K, B, W, name) in [(10, 3, 64, 20, "C3 (K3,N10,B64)"), (20, 3, 64, 28, "C5 shard (K3,N20,B64/GPU)")]:
    sd = po.init_state_dict(K, seed=1)
    x, S = synthetic.make_batch(B, N, W, seed=3)
    tgt = torch.from_numpy(synthetic.random_targets(B, N, seed=4))
    xt, St, tt = torch.from_numpy(x).cuda(), torch.from_numpy(S).cuda(), tgt.cuda()
    for path in ("native",):
        m = gp.DecentralPlannerNet(Cfg(N, K)); m.load_state_dict(sd); m = m.cuda().train()
        opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5)

        def step():
            opt.zero_grad(set_to_none=True)
            m.addGSO(St)
            loss = gp.planner_loss(m.forward_logits(xt), tt)
            loss.backward()
            opt.step()
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(30):
            step()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 30
        print("%-28s %-7s %8.3f ms/step  %8.1f K agent-steps/s" % (name, path, ms, B * N / ms))
    # CPU oracle step (reference's op sequence), 8 threads max
    torch.set_num_threads(min(16, os.cpu_count() or 1))
    names = [k for k, v in sd.items() if v.is_floating_point() and "running" not in k]
    leaf = {k: sd[k].clone().requires_grad_(True) for k in names}
    full = dict(sd); full.update(leaf)
    bn = {k: v.clone() for k, v in sd.items() if "running" in k or "tracked" in k}
    optc = torch.optim.Adam([leaf[k] for k in names], lr=1e-3, weight_decay=1e-5)
    xc, Sc = torch.from_numpy(x), torch.from_numpy(S)
    def cstep():
        optc.zero_grad()
        l = po.planner_loss(po.planner_forward(full, Sc, xc, True, bn), tgt); l.backward(); optc.step()
    cstep()
    t0 = time.perf_counter(); n = 0
    while time.perf_counter() - t0 < 4.0:
        cstep(); n += 1
    ms = (time.perf_counter() - t0) / n * 1e3
    print("%-28s %-7s %8.3f ms/step  %8.1f K agent-steps/s (CPU oracle, %d threads)" % (name, "cpu", ms, B * N / ms, torch.get_num_threads()))
