#!/usr/bin/env python
"""Micro-benchmark of the field kernels (GB/s vs measured HBM copy bandwidth)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from debvader_b200 import _fieldops

def timeit(fn, iters=10):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / iters

F, S, C = 4096, 59, 6
dev = torch.device("cuda")
field = (torch.randn((1, F, F, C), device=dev) * 0.6).double()
rng = np.random.default_rng(5)
# reference: plain device copy of the same number of bytes
n = 16384 * S * S * C
src = torch.randn(n, device=dev, dtype=torch.float64); dst = torch.empty_like(src)
t = timeit(lambda: dst.copy_(src)); print(f"torch copy f64 {2*n*8/t/1e6:.0f} GB/s")
for N in (2000, 16384):
    c = rng.integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2)).astype(np.float64)
    plan = _fieldops.plan_windows(c, S, F)
    for od, bps in ((torch.float64, 16), (torch.float32, 12)):
        t = timeit(lambda: _fieldops.extract(field, plan, S, C, out_dtype=od))
        print(f"extract N={N} f64->{'f64' if bps==16 else 'f32'}: {t:.3f} ms  {N*S*S*C*bps/t/1e6:.0f} GB/s (PER={os.environ.get('DBV_EXTRACT_PER','auto')})")
N = 2000
c = rng.integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2))
st = torch.randn((N, S, S, C), device=dev)
off = _fieldops.subtract_offset(F, S)
res = torch.empty_like(field)
t = timeit(lambda: _fieldops.window_axpy(field, st, off + c[:, 0], off + c[:, 1], -1.0, out=res))
moved = 2 * field.numel() * 8 + N * S * S * C * 4
print(f"window_axpy N={N}: {t:.3f} ms  moved {moved/t/1e6:.0f} GB/s, algorithmic {N*S*S*C*20/t/1e6:.0f} GB/s")
xd = torch.from_numpy((off + c[:, 0]).astype(np.int32)).to(dev); yd = torch.from_numpy((off + c[:, 1]).astype(np.int32)).to(dev)
t = timeit(lambda: _fieldops.window_axpy(field, st, xd, yd, -1.0, out=res))
print(f"window_axpy N={N} (positions resident): {t:.3f} ms  moved {moved/t/1e6:.0f} GB/s, algorithmic {N*S*S*C*20/t/1e6:.0f} GB/s")
work = field.clone()
t = timeit(lambda: _fieldops.window_axpy(work, st, xd, yd, -1.0, out=work))
print(f"window_axpy N={N} in place (positions resident): {t:.3f} ms  algorithmic {N*S*S*C*20/t/1e6:.0f} GB/s")
t = timeit(lambda: _fieldops.window_axpy(field, st[:0], xd[:0], yd[:0], -1.0, out=res))
print(f"window_axpy N=0 (plain tiled copy): {t:.3f} ms  {2*field.numel()*8/t/1e6:.0f} GB/s")
t = timeit(lambda: res.copy_(field)); print(f"torch copy of the field alone: {t:.3f} ms {2*field.numel()*8/t/1e6:.0f} GB/s")
a = torch.randn_like(field)
t = timeit(lambda: _fieldops.mse(field, a)); print(f"field mse: {t:.3f} ms {2*field.numel()*8/t/1e6:.0f} GB/s")
# ---- sub-pixel placement: 2000 stamps at fractional positions on the 4096^2 field ------------------
fpos = c + rng.uniform(-0.5, 0.5, size=c.shape)
E = _fieldops.spline_extent(S)
t = timeit(lambda: _fieldops.spline_place(st[:512], fpos[:512, 0], fpos[:512, 1], F), iters=5)
print(f"spline_place 512 stamps -> ({E},{E}) f64 windows: {t:.3f} ms  ({512*(S*S*C*4 + E*E*C*8)/t/1e6:.0f} GB/s of stamp read + window write)")
t = timeit(lambda: _fieldops.spline_window_axpy(field, st, fpos[:, 0], fpos[:, 1], -1.0), iters=3)
print(f"sub-pixel residual of {N} stamps (place + paste, one batch): {t:.3f} ms")
t = timeit(lambda: _fieldops.spline_window_axpy(field, st, fpos[:, 0], fpos[:, 1], -1.0, batch=512), iters=3)
print(f"sub-pixel residual of {N} stamps (place + paste, batches of 512): {t:.3f} ms")
# ---- position fit ---------------------------------------------------------------------------------
import time
from debvader_b200.deblend_cutout.optimization import FieldBand, fit_position
fb = FieldBand(field)
blob = torch.zeros((S, S), device=dev, dtype=torch.float64); blob[24:35, 24:35] = 5.0
torch.cuda.synchronize(); t0 = time.perf_counter(); nfev = 0
for k in range(20):
    r = fit_position(fb, blob, c[k].astype(np.float64), return_result=True); nfev += r.nfev
torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"position fit: {dt/20*1e3:.2f} ms per galaxy, {nfev/20:.1f} least_squares nfev per galaxy (each with 2-point Jacobian evaluations)")
