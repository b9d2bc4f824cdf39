"""Whole-planner inference (device-resident, CUDA events) for the BASELINE.json inference configs, per feature-
extractor kernel.  usage: python profiles/planner_microbench.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import synthetic
from oracle import planner_oracle as po


class Cfg:
    def __init__(self, n, k):
        self.num_agents, self.nGraphFilterTaps, self.device = n, k, torch.device("cuda")


for (N, K, B, W, name) in [(10, 3, 64, 20, "C2"), (40, 3, 256, 50, "C4"), (20, 3, 512, 28, "C5-shape inference"),
                           (10, 3, 4096, 20, "large")]:
    sd = po.init_state_dict(K, seed=1)
    po.randomize_bn_stats(sd)
    m = gp.DecentralPlannerNet(Cfg(N, K))
    m.load_state_dict(sd)
    m = m.cuda().eval()
    x, S = synthetic.make_batch(min(B, 64), N, W, seed=3)
    reps = (B + x.shape[0] - 1) // x.shape[0]
    xt = torch.from_numpy(x).repeat(reps, 1, 1, 1, 1)[:B].cuda()
    St = torch.from_numpy(S).repeat(reps, 1, 1)[:B].cuda()
    for fe in ("cuda", "tc"):
        for gf in ("cuda", "tc"):
            m.set_feature_mode(fe)
            m.set_graph_filter_mode(gf)
            with torch.no_grad():
                for _ in range(5):
                    m.addGSO(St); out = m(xt)
                torch.cuda.synchronize()
                iters = 200 if B * N < 20000 else 30
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    m.addGSO(St); out = m(xt)
                e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / iters
            print("%-20s N=%2d B=%5d  feature=%-4s filter=%-4s  %9.1f us/step  %8.2f M agent-steps/s"
                  % (name, N, B, fe, gf, us, B * N / us))
