#!/usr/bin/env python
"""cProfile of one DeblendField API pass (4096^2 field, 2000 sources): where the host time of `ms_per_field` goes."""
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

F, N = 4096, 2000
dev = torch.device("cuda")
field = (torch.randn((1, F, F, 6), device=dev) * 0.6).double()
centres = np.random.default_rng(5).integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2)).astype(np.float64)
net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")
obj = DeblendField(net, field)


def one():
    obj.deblend_field(centres)
    r = obj.get_residual_field(as_tensor=True)
    return obj.field_mse(obj.field_tensor, r)


with contextlib.redirect_stdout(io.StringIO()):
    one()
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    pr.disable()
st = io.StringIO()
pstats.Stats(pr, stream=st).sort_stats("cumulative").print_stats(35)
print(st.getvalue())
