#!/usr/bin/env python
"""Does running two contexts on two streams (each on part of the batch) fill the tails / prologues of the persistent kernels?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from debvader_b200.model.model import load_deblender
CFG = ("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3])
B = 4096
x = torch.randn((B, 59, 59, 6), device="cuda"); m = torch.empty_like(x); s = torch.empty_like(x)
def t(fn, it=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(it): fn()
    b.record(); torch.cuda.synchronize(); return a.elapsed_time(b) / it
net = load_deblender(*CFG, weights="random:1234", chunk=4096)
print("one ctx, chunk 4096: %.2f ms" % t(lambda: net.deblend_into(x, m, s)))
net.close()
for piece in (2048, 1024):
    nets = [load_deblender(*CFG, weights="random:1234", chunk=piece) for _ in range(2)]
    print("one ctx, %d-stamp calls: %.2f ms" % (piece, t(lambda: [nets[0].deblend_into(x[b:b + piece], m[b:b + piece], s[b:b + piece]) for b in range(0, B, piece)])))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    def run():
        cur = torch.cuda.current_stream()
        for st in streams: st.wait_stream(cur)
        for i, b in enumerate(range(0, B, piece)):
            with torch.cuda.stream(streams[i & 1]):
                nets[i & 1].deblend_into(x[b:b + piece], m[b:b + piece], s[b:b + piece])
        for st in streams: cur.wait_stream(st)
    print("two ctxs on two streams, %d-stamp calls alternating: %.2f ms" % (piece, t(run)))
    for n in nets: n.close()
