import os, sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from debvader_b200.model.model import load_deblender
from debvader_b200.deblend_cutout.deblender import deblend
net = load_deblender("dc2", (59,59,6), 32, [32,64,128,256], [3,3,3,3], weights="random:1234")
B = 4096
x_host = torch.empty((B,59,59,6), dtype=torch.float32, pin_memory=True); x_host.normal_()
xh = x_host.numpy()
for _ in range(3): deblend(net, xh)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(10): deblend(net, xh)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 10
print(os.environ.get("DBV_HOST_PIECE"), "e2e ms", round(dt*1e3, 2), "stamps/s", round(B/dt))
