"""Per-role cycle accounting of gf_fwd_pair_kernel ("tc_timing" debug option).  usage: python profiles/pair_phase_timing.py [B]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import gnn_pathplanning_b200 as gp
from gnn_pathplanning_b200 import _lib

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N, K = 10, 3
torch.manual_seed(0)
w = ((torch.rand(128, 1, K, 128) - 0.5) * 0.2).cuda()
b = (torch.rand(128, 1) - 0.5).cuda()
x = torch.randn(B, N, 128, device="cuda")
S = torch.rand(B, N, N, device="cuda") * 0.2
lib = _lib.load()
_lib.set_debug_option("gf_mode", 3)
for _ in range(3):
    y = gp.graph_filter(x, S, w, b, True, gp.NODE_MAJOR, gp.NODE_MAJOR)
_lib.set_debug_option("tc_timing", 1)
y = gp.graph_filter(x, S, w, b, True, gp.NODE_MAJOR, gp.NODE_MAJOR)
out = (C.c_ulonglong * 16)()
lib.gpp_debug_pair_timing(out)
reps = 5
for _ in range(reps):
    y = gp.graph_filter(x, S, w, b, True, gp.NODE_MAJOR, gp.NODE_MAJOR)
lib.gpp_debug_pair_timing(out)
v = [float(o) for o in out]
ctas, pairs = v[14], v[11]
tiles_per_cta = 2 * pairs / ctas
print("B=%d: %d CTAs, %.1f tiles per CTA per launch, kernel %.0f cycles per CTA" % (B, ctas / reps, tiles_per_cta, v[12] / ctas))
per_tile = lambda c, n: c / max(n, 1)
nt_cta = pairs * 2          # tiles over all CTAs
print("per tile (cycles): producer warp 0: wait S %.0f, wait ring %.0f, item work %.0f (%.2f items, %.0f cycles per item)"
      % (v[0] / nt_cta, v[1] / nt_cta, v[2] / nt_cta, v[3] / nt_cta, v[2] / max(v[3], 1)))
print("                   MMA thread (per tile pair): wait operands %.0f, wait accumulator %.0f, loop %.0f; start-up %.0f"
      % (v[4] / pairs, v[5] / pairs, v[6] / pairs, v[13] / (ctas / 2)))
print("                   epilogue warp 0: wait %.0f, work %.0f;  scout: wait %.0f, work %.0f"
      % (v[7] / nt_cta, v[8] / nt_cta, v[9] / nt_cta, v[10] / nt_cta))
