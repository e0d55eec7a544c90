#!/usr/bin/env python
"""three calls of the sub-pixel placement (512 stamps) for `ncu`: per-kernel durations of weights / pass X / pass Y / paste"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from debvader_b200 import _fieldops
F, S, C, N = 4096, 59, 6, 512
rng = np.random.default_rng(5)
dev = torch.device("cuda")
field = (torch.randn((1, F, F, C), device=dev) * 0.6).double()
st = torch.randn((N, S, S, C), device=dev)
pos = rng.integers(-(F // 2 - 70), F // 2 - 70, size=(N, 2)) + rng.uniform(-0.5, 0.5, size=(N, 2))
for _ in range(3):
    placed, ax, ay = _fieldops.spline_place(st, pos[:, 0], pos[:, 1], F)
    out = _fieldops.window_axpy(field, placed, ax, ay, -1.0, planar=True)
torch.cuda.synchronize()
print("ok", float(out.sum()))
