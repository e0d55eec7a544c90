#!/usr/bin/env python
"""Where the end-to-end time of deblend(net, host array) goes: API wrapper vs the C pipeline vs compute at piece granularity."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from debvader_b200.model.model import load_deblender
from debvader_b200.deblend_cutout.deblender import deblend
net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")
B = 4096
x_host = torch.empty((B, 59, 59, 6), dtype=torch.float32, pin_memory=True); x_host.normal_()
xh = x_host.numpy()
out = torch.empty((B, 59, 59, 6), dtype=torch.float32, pin_memory=True).numpy()
def t(fn, it=8):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(it): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / it * 1e3
print("deblend(net, host)                         %.2f ms" % t(lambda: deblend(net, xh)))
print("deblend_host(resident=True)                %.2f ms" % t(lambda: net.deblend_host(xh, resident=True)))
print("deblend_host(out_mean=pre, no stddev)      %.2f ms" % t(lambda: net.deblend_host(xh, want_stddev=False, out_mean=out)))
xd = x_host.cuda(); m = torch.empty_like(xd); s = torch.empty_like(xd)
print("device, one call of 4096                   %.2f ms" % t(lambda: net.deblend_into(xd, m, s)))
for piece in (512, 1024, 2048):
    def run():
        for b0 in range(0, B, piece): net.deblend_into(xd[b0:b0 + piece], m[b0:b0 + piece], s[b0:b0 + piece])
    print("device, %4d-stamp calls                    %.2f ms" % (piece, t(run)))
print("growing-piece schedule (default)           %.2f ms" % t(lambda: net.deblend_host(xh, want_stddev=False, out_mean=out)))
for piece in ("1024",):
    os.environ["DBV_HOST_PIECE"] = piece
    print("DBV_HOST_PIECE=%-5s deblend_host(pre)       %.2f ms" % (piece, t(lambda: net.deblend_host(xh, want_stddev=False, out_mean=out))))
