#!/usr/bin/env python
"""BASELINE config 4 on N GPUs of one box: one deblending pass over a synthetic field, tiled across ranks
(debvader_b200.parallel.deblend_field_tiled), checked bit for bit against the single-GPU result on rank 0.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/field_tiled_nccl.py [F] [sources]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from debvader_b200 import _fieldops, parallel as par  # noqa: E402
from debvader_b200.model.model import load_deblender  # noqa: E402


def main():
    F_ = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    S, C = 59, 6
    g = torch.Generator(device=dev).manual_seed(5)  # same field on every rank
    field = (torch.randn((1, F_, F_, C), device=dev, generator=g, dtype=torch.float32) * 0.6).double()
    rng = np.random.default_rng(5)
    centres = rng.integers(-(F_ // 2 - 30), F_ // 2 - 30, size=(N, 2)).astype(np.float64)
    net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")

    def one():
        return par.deblend_field_tiled(net, field, centres, sample=False)

    tile, (r0, r1, c0, c1), idx = one()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(3):
        tile, _, _ = one()
    torch.cuda.synchronize()
    dt = torch.tensor([(time.perf_counter() - t0) / 3 * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    # single-GPU truth of this rank's tile (every rank can compute it: it holds the field)
    plan = _fieldops.plan_windows(centres, S, F_)
    cut, lidx = _fieldops.extract(field, plan, S, C, out_dtype=torch.float32)
    assert lidx == idx
    mean = net(cut, sample=False).mean().tensor
    off = _fieldops.subtract_offset(F_, S)
    full = _fieldops.window_axpy(field, mean, off + centres[:, 0].astype(np.int64), off + centres[:, 1].astype(np.int64), -1.0)
    same = bool(torch.equal(tile, full[0, r0:r1, c0:c1]))
    flag = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"config": f"one deblending pass, {F_}x{F_}x6 f64 field, {N} sources, tiled over {world} GPU(s) ({par.tile_grid(world)} owner tiles)",
                          "ms_per_field_pass": float(dt.item()), "tiles_bit_identical_to_single_gpu": bool(flag.item()), "n_gpus": world}), flush=True)
    assert bool(flag.item()), "tiled residual differs from the single-GPU residual"
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
