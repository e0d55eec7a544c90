#!/usr/bin/env python
"""Raw host<->device copy rates of the box (pinned memory): what bounds the end-to-end number.

    python tools/pcie_probe.py                                   one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/pcie_probe.py
                                                                 N ranks copying at the same time (the aggregate ceiling
                                                                 of the box, next to the end-to-end rate bench.py reaches)
One JSON line (rank 0): per-rank and aggregate GB/s for H2D alone, D2H alone and both directions at once, and the
stamps/s ceiling these imply for `deblend(net, host array)` (4096 stamps in, 4096 means out per step)."""
import json
import os
import time

import torch

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
n = 4096 * 59 * 59 * 6
h_in = torch.empty(n, dtype=torch.float32, pin_memory=True).normal_()
h_out = torch.empty(n, dtype=torch.float32, pin_memory=True)
d_in = torch.empty(n, dtype=torch.float32, device=dev)
d_out = torch.randn(n, dtype=torch.float32, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def t(fn, it=10):
    fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(it):
        fn()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / it], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)  # the slowest rank: all ranks copy concurrently
    return float(dt.item())


def h2d():
    with torch.cuda.stream(s1):
        d_in.copy_(h_in, non_blocking=True)


def d2h():
    with torch.cuda.stream(s2):
        h_out.copy_(d_out, non_blocking=True)


def both():
    h2d()
    d2h()


gb = n * 4 / 1e9
a, b, c = t(h2d), t(d2h), t(both)
if rank == 0:
    out = {"n_gpus": world, "bytes_per_direction_per_rank": n * 4,
           "h2d_GBps_per_rank": gb / a, "d2h_GBps_per_rank": gb / b, "both_ms_per_4096_stamps": c * 1e3,
           "aggregate_GBps_both_directions": world * 2 * gb / c,
           "e2e_ceiling_stamps_per_s": world * 4096 / c,
           "note": "pinned buffers, one copy stream per direction, all ranks copying at the same time, max over ranks"}
    print(json.dumps(out))
    print(f"H2D {gb/a:.1f} GB/s ({a*1e3:.2f} ms per 4096 stamps)  D2H {gb/b:.1f} GB/s ({b*1e3:.2f} ms)  both at once {c*1e3:.2f} ms -> "
          f"{world*4096/c:.0f} stamps/s ceiling of the e2e path on {world} GPU(s)")
if world > 1:
    dist.destroy_process_group()
