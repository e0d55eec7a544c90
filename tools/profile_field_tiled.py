#!/usr/bin/env python
"""cProfile of one TILED DeblendField pass on one rank with the per-rank load of an 8-GPU run (an eighth of the sources,
an eighth of the field): where the host time goes once the GPU work per rank is small."""
import cProfile
import contextlib
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from debvader_b200.deblend.field_deblender import DeblendField
from debvader_b200.model.model import load_deblender

F, N = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1536, 250)
field = np.random.default_rng(5).standard_normal((1, F, F, 6)) * 0.6
centres = np.random.default_rng(6).integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2)).astype(np.float64)
net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")
net.sample = False
obj = DeblendField(net, field, tiled=True)


def one():
    obj.deblend_field(centres)
    r = obj.get_residual_field(as_tensor=True)
    return obj.field_mse(obj.field_tensor, r)


with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        one()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(10):
        one()
    torch.cuda.synchronize()
    pr.disable()
print(f"{F}^2 field, {N} sources, one rank: {ms:.3f} ms per pass")
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("tottime").print_stats(28)
print(st.getvalue())
