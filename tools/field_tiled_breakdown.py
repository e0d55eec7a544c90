#!/usr/bin/env python
"""Where the time of a tiled field pass goes on N ranks: every phase of DeblendField(tiled=True) bracketed by a device
synchronize + host clock, per rank (max and mean over ranks printed by rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/field_tiled_breakdown.py [F] [sources]
"""
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from debvader_b200 import parallel as par  # noqa: E402
from debvader_b200.deblend.field_deblender import DeblendField  # noqa: E402
from debvader_b200.model.model import load_deblender  # noqa: E402

F_ = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
field = np.random.default_rng(5).standard_normal((1, F_, F_, 6), dtype=np.float32).astype(np.float64) * 0.6
centres = np.random.default_rng(5).integers(-(F_ // 2 - 30), F_ // 2 - 30, size=(N, 2)).astype(np.float64)
net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")
net.sample = False
obj = DeblendField(net, field, tiled=True)
T = {}


def lap(name, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    T[name] = T.get(name, 0.0) + (t1 - t0) * 1e3
    return t1


def one(timed):
    t = time.perf_counter()
    obj.deblend_field(centres)
    if timed:
        t = lap("deblend_field", t)
    r = obj.get_residual_field(as_tensor=True)
    if timed:
        t = lap("get_residual_field (exchange + subtract)", t)
    m = obj.field_mse(obj.field_tensor, r)
    if timed:
        t = lap("field_mse (partial sum + all-reduce)", t)
    return m


iters = 10
with contextlib.redirect_stdout(io.StringIO()):
    for _ in range(3):
        one(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(iters):
        one(False)
    torch.cuda.synchronize()
    untimed = (time.perf_counter() - t0) * 1e3 / iters
    if world > 1:
        dist.barrier()
    for _ in range(iters):
        one(True)
names = list(T)
v = torch.tensor([untimed] + [T[k] / iters for k in names], device=dev, dtype=torch.float64)
mx, mean = v.clone(), v.clone()
if world > 1:
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(mean, op=dist.ReduceOp.SUM)
    mean /= world
if rank == 0:
    out = {"n_gpus": world, "field": F_, "sources": N, "ms_per_pass_free_running": {"max": mx[0].item(), "mean": mean[0].item()},
           "phases_with_a_synchronize_after_each": {k: {"max": mx[i + 1].item(), "mean": mean[i + 1].item()} for i, k in enumerate(names)},
           "sources_owned_rank0": int(len(obj._tile_state[2])) if obj._tile_state else None}
    print(json.dumps(out))
if world > 1:
    dist.destroy_process_group()
