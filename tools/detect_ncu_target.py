#!/usr/bin/env python
"""The device detector on a 4096^2 x 6 f64 field with 2000 round blobs (bench.py's `detect` extra), three calls: the target of
the `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` pass committed under profiles/
(per-kernel times / DRAM bytes of csrc/detect_kernels.cu).  Without ncu it prints CUDA-event times."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from debvader_b200.detect.detection import DeviceDetector

F, C, N = 4096, 6, 2000
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(6)
field = (torch.randn((1, F, F, C), device=dev, generator=g, dtype=torch.float32) * 0.03).double()
rng = np.random.default_rng(5)
yy, xx = np.mgrid[-15:16, -15:16]
for (px, py) in rng.uniform(40, F - 40, (N, 2)):
    ix, iy = int(px), int(py)
    blob = rng.uniform(0.5, 3.0) * np.exp(-((xx - (px - ix)) ** 2 + (yy - (py - iy)) ** 2) / (2 * rng.uniform(1.2, 2.5) ** 2))
    field[0, iy - 15 : iy + 16, ix - 15 : ix + 16, :] += torch.from_numpy(blob).to(dev)[..., None]
d = DeviceDetector(device=dev)
c = d(field)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    d.run(field)
b.record()
torch.cuda.synchronize()
t0 = time.perf_counter()
c = d(field)
t1 = time.perf_counter()
print({"objects": int(len(c)), "ms_enqueued": a.elapsed_time(b) / 3, "ms_call_with_centres_on_host": (t1 - t0) * 1e3})
