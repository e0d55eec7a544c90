#!/usr/bin/env python
"""Raw host<->device copy rates of the box (pinned memory): what bounds the end-to-end number."""
import time, torch
n = 4096 * 59 * 59 * 6
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device="cuda")
d_out = torch.randn(n, dtype=torch.float32, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, it=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
def both():
    h2d(); d2h()
gb = n * 4 / 1e9
a, b, c = t(h2d), t(d2h), t(both)
print(f"H2D {gb/a:.1f} GB/s ({a*1e3:.2f} ms per 4096 stamps)  D2H {gb/b:.1f} GB/s ({b*1e3:.2f} ms)  both at once {c*1e3:.2f} ms -> {4096/c:.0f} stamps/s ceiling of the e2e path")
