#!/usr/bin/env python
"""End-to-end rate of deblend(net, host array) for pageable float64 / float32 input against the number of staging threads
(DEBVADER_B200_HOST_THREADS): python tools/e2e_pageable.py"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r"""
import sys, time, numpy as np, torch
sys.path.insert(0, %r)
from debvader_b200.model.model import load_deblender
from debvader_b200.deblend_cutout.deblender import deblend
net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")
x = (np.random.default_rng(0).standard_normal((4096, 59, 59, 6)) * 0.3)
for name, a in (("pageable f64", x), ("pageable f32", x.astype(np.float32))):
    for _ in range(2): deblend(net, a)
    t0 = time.perf_counter()
    for _ in range(5): deblend(net, a)
    dt = (time.perf_counter() - t0) / 5
    print(name, round(4096 / dt), "stamps/s", round(dt * 1e3, 2), "ms")
""" % ROOT
for th in sys.argv[1:] or ["2", "4", "8", "12", "16"]:
    env = dict(os.environ, DEBVADER_B200_HOST_THREADS=th)
    out = subprocess.run([sys.executable, "-c", CODE], env=env, capture_output=True, text=True)
    print(f"threads={th}:", " | ".join(l for l in out.stdout.splitlines() if "stamps/s" in l) or out.stderr[-300:])
