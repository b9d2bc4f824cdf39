"""Runs the standalone node-major graph filter a few times (for ncu captures).
usage: python profiles/run_gf_once.py [B] [N] [K] [gf_mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gnn_pathplanning_b200 as gp

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
N = int(sys.argv[2]) if len(sys.argv) > 2 else 10
K = int(sys.argv[3]) if len(sys.argv) > 3 else 3
from gnn_pathplanning_b200 import _lib
_lib.set_debug_option("gf_mode", int(sys.argv[4]) if len(sys.argv) > 4 else 0)
torch.manual_seed(0)
w = ((torch.rand(128, 1, K, 128) - 0.5) * 0.2).cuda()
b = (torch.rand(128, 1) - 0.5).cuda()
x = torch.randn(B, N, 128, device="cuda")
S = torch.rand(B, N, N, device="cuda") * 0.2
for _ in range(4):
    y = gp.graph_filter(x, S, w, b, True, gp.NODE_MAJOR, gp.NODE_MAJOR)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
