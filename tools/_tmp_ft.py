import sys, os, json, time, gc, io, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from debvader_b200.model.model import load_deblender
from debvader_b200.deblend.field_deblender import DeblendField
dev = torch.device("cuda", 0); torch.cuda.set_device(0)
net = load_deblender(*bench.CFG, weights="random:1234", precision="mixed", chunk=4096)
net.sample = False
F, N = 4096, 2000
field = np.random.default_rng(5).standard_normal((1, F, F, 6), dtype=np.float32).astype(np.float64) * 0.6
centres = np.random.default_rng(6).integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2)).astype(np.float64)
gcs = []
def cb(phase, info):
    if phase == "start": cb.t = time.perf_counter()
    else: gcs.append((info["generation"], (time.perf_counter() - cb.t) * 1e3))
gc.callbacks.append(cb)
for tiled in (True, False):
    obj = DeblendField(net, field if tiled else torch.from_numpy(field).to(dev), tiled=tiled)
    keep = {}
    def one():
        t0 = time.perf_counter()
        obj.deblend_field(centres); t1 = time.perf_counter()
        keep["res"] = obj.get_residual_field(as_tensor=True); t2 = time.perf_counter()
        keep["mse"] = obj.field_mse(obj.field_tensor, keep["res"]); t3 = time.perf_counter()
        return [(t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3]
    with contextlib.redirect_stdout(io.StringIO()):
        one(); torch.cuda.synchronize()
        rows = []
        for _ in range(12):
            gcs.clear()
            r = one()
            rows.append((r, list(gcs)))
    print("tiled" if tiled else "single", "host ms per phase [deblend_field, residual, mse] + GCs (gen, ms):")
    for r, g in rows: print("   ", [round(v, 2) for v in r], [(a, round(b, 2)) for a, b in g if b > 0.2])
