"""Per-call device time of one lock-step rollout step (inputs builder, planner forward, move).
usage: python profiles/rollout_breakdown.py [episodes] [agents] [map_w]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import synthetic
from oracle import planner_oracle as po


class Cfg:
    def __init__(self, n, k):
        self.num_agents, self.nGraphFilterTaps, self.device = n, k, torch.device("cuda")


E = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
W = int(sys.argv[3]) if len(sys.argv) > 3 else 20
rng = np.random.default_rng(1)
cases = [synthetic.random_episode(rng, N, W, 0.1) for _ in range(min(E, 64))]
rep = (E + len(cases) - 1) // len(cases)
maps = np.stack([c[0] for c in cases] * rep)[:E]; starts = np.stack([c[1] for c in cases] * rep)[:E]
goals = np.stack([c[2] for c in cases] * rep)[:E]
sd = po.init_state_dict(3, seed=1); sd["actionsMLP.0.weight"] = sd["actionsMLP.0.weight"] * 40.0
m = gp.DecentralPlannerNet(Cfg(N, 3)); m.load_state_dict(sd); m = m.cuda().eval()
ro = gp.BatchedRollout(N, 6.0).setup(starts, goals, maps, 1 << 30)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
acc = np.zeros(3)
with torch.no_grad():
    for i in range(60):
        ev[0].record()
        x, S = ro.build_inputs(i + 1)
        ev[1].record()
        m.addGSO(S)
        lg = m.forward_logits(x)
        ev[2].record()
        ro.move(lg, i + 2)
        ev[3].record()
        torch.cuda.synchronize()
        if i >= 10:
            acc += [ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])]
acc *= 1e3 / 50
print("episodes %d x %d agents, %dx%d map: inputs %.1f us  planner forward %.1f us  move %.1f us  (events around each call, "
      "synchronised per step)" % (E, N, W, W, acc[0], acc[1], acc[2]))
