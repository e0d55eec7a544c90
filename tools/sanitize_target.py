#!/usr/bin/env python
"""Small invocations of every kernel family for `compute-sanitizer --tool memcheck` / `racecheck`:

    compute-sanitizer --tool memcheck --error-exitcode 1 python tools/sanitize_target.py [precision ...]

The context is created with a small chunk so that the plan autotuner and the activations stay small; the field operators
run on a 200^2 field."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from debvader_b200 import _fieldops
from debvader_b200.model.model import load_deblender

precisions = sys.argv[1:] or ["mixed", "fp32tc", "bf16x3", "fp32"]
x = torch.randn((70, 59, 59, 6), device="cuda") * 0.3
for p in precisions:
    net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234", precision=p, chunk=64)
    d = net(x, sample=False)  # two chunks: 64 + 6 stamps
    torch.cuda.synchronize()
    assert torch.isfinite(d.mean().tensor).all()
    net.close()
    print("network", p, "ok")
rng = np.random.default_rng(0)
F, S, C, N = 200, 59, 6, 90
field = torch.from_numpy(rng.normal(size=(1, F, F, C))).cuda()
centres = rng.integers(-F // 2 - 10, F // 2 + 10, size=(N, 2)).astype(np.float64)
plan = _fieldops.plan_windows(centres, S, F)
for dt in (torch.float64, torch.float32):
    cut, idx = _fieldops.extract(field, plan, S, C, out_dtype=dt)
stamps = torch.randn((N, S, S, C), device="cuda")
off = _fieldops.subtract_offset(F, S)
x0, y0 = off + centres[:, 0].astype(int), off + centres[:, 1].astype(int)
res = _fieldops.window_axpy(field, stamps, x0, y0, -1.0)
work = field.clone()
_fieldops.window_axpy(work, stamps, x0, y0, -1.0, out=work)
assert torch.equal(work, res)
_fieldops.window_axpy(None, stamps[:, :, :, :2].contiguous(), x0, y0, 1.0, field_shape=(F, F, 2))  # generic kernel (C = 2)
_fieldops.mse(field, res)
_fieldops.center_mse(cut.double(), stamps, 24, 34)
_fieldops.spline_window_axpy(field, stamps[:8], centres[:8, 0] + 0.3, centres[:8, 1] - 0.2, -1.0)
torch.cuda.synchronize()
print("field operators ok")
