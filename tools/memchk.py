import sys, time
sys.path.insert(0, "/root/repo")
import torch
from debvader_b200.model.model import load_deblender
torch.cuda.init()
f0, t = torch.cuda.mem_get_info()
t0 = time.perf_counter()
net = load_deblender("dc2", (59,59,6), 32, [32,64,128,256], [3,3,3,3], weights="random:1234")
torch.cuda.synchronize()
dt = time.perf_counter() - t0
f1, _ = torch.cuda.mem_get_info()
print(f"load_deblender: {dt:.2f} s, device memory used by the context: {(f0 - f1) / 2**30:.2f} GiB of {t / 2**30:.0f} GiB")
