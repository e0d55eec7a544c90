#!/usr/bin/env python
"""BASELINE config 4 on N GPUs of one box: deblending passes over a synthetic field tiled across ranks — every rank
holds ONLY its owner tile + 30-px halo (debvader_b200.parallel.LocalField) — through the public API
(DeblendField(tiled=True).deblend_field + get_residual_field(as_tensor=True) + field_mse), checked bit for bit against
the single-GPU result (computed by rank 0 alone on the full field and broadcast as per-region checksums), followed by the
iterative loop (IterativeDeblendField(tiled=True)) with a given list of centres per step.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/field_tiled_nccl.py [F] [sources]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from debvader_b200 import parallel as par  # noqa: E402
from debvader_b200.deblend.field_deblender import DeblendField  # noqa: E402
from debvader_b200.deblend_iterative.iterative_deblender import IterativeDeblendField  # noqa: E402
from debvader_b200.model.model import load_deblender  # noqa: E402


def synthetic_field(F_, C, seed=5):
    rng = np.random.default_rng(seed)
    return rng.standard_normal((1, F_, F_, C), dtype=np.float32).astype(np.float64) * 0.6


def field_pass_tiled(net, field_host_or_local, centres, iters=3):
    """ms per (deblend_field + residual + field MSE) pass on a tiled field, CUDA events, max over ranks taken by the caller."""
    obj = DeblendField(net, field_host_or_local, tiled=True)
    out = {}

    def one():
        obj.deblend_field(centres)
        res = obj.get_residual_field(as_tensor=True)
        out["mse"] = obj.field_mse(obj.field_tensor, res)
        out["res"] = res

    one()
    one()  # steady state alternates between two result buffers; the first cudaMalloc of a field-sized buffer costs ~100 ms
    torch.cuda.synchronize()
    if dist.is_initialized():
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        one()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters, obj, out


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
    field = synthetic_field(F_, C)  # the same host array in every process; only the local region goes to the GPU
    rng = np.random.default_rng(5)
    centres = rng.integers(-(F_ // 2 - 30), F_ // 2 - 30, size=(N, 2)).astype(np.float64)
    raw = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234")
    raw.sample = False  # z = loc: the pass is deterministic, so tiles can be compared bit for bit

    torch.cuda.reset_peak_memory_stats()
    base_mem = torch.cuda.memory_allocated()
    ms, obj, out = field_pass_tiled(raw, field, centres)
    dt = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    loc = obj._local
    R0, R1, C0, C1 = loc.region
    share = loc.nbytes() / (F_ * F_ * C * 8)

    # single-GPU truth: rank 0 runs the whole field alone, then every rank compares ITS region
    # (the truth is sent region by region so that no rank ever holds the full field on its GPU but rank 0)
    same = True
    mse_single = None
    if rank == 0:
        single = DeblendField(raw, torch.from_numpy(field).to(dev))
        single.deblend_field(centres)
        full = single.get_residual_field(as_tensor=True)
        mse_single = single.field_mse(single.field_tensor, full)
        regions = par.region_bounds(F_, world)
        for r, (a0, a1, b0, b1) in enumerate(regions):
            part = full[:, a0:a1, b0:b1].contiguous()
            if r == 0:
                same = bool(torch.equal(part, out["res"]))
            else:
                dist.send(part, dst=r)
        del full, single
    elif world > 1:
        truth = torch.empty_like(out["res"])
        dist.recv(truth, src=0)
        same = bool(torch.equal(truth, out["res"]))
    flag = torch.tensor([1 if same else 0], device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    # iterative loop (cfg 4): three steps with given centres (detection excluded)
    steps = [centres[: N // 2], centres, centres[: N // 4]]
    calls = []

    def detector(region, local_field):
        calls.append(1)
        return steps[min(len(calls) - 1, len(steps) - 1)]

    detector.accepts_local = True
    it = IterativeDeblendField(raw, loc, detector=detector, tiled=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    import contextlib
    import io

    with contextlib.redirect_stdout(io.StringIO()):
        it.iterative_deblending()
    b.record()
    torch.cuda.synchronize()
    it_ms = torch.tensor([a.elapsed_time(b)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(it_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({
            "config": f"deblend_field + get_residual_field + field MSE through DeblendField(tiled=True), {F_}x{F_}x6 f64 field, {N} sources, "
                      f"tiled over {world} GPU(s) ({par.tile_grid(world)} owner tiles + {par.HALO}-px halo)",
            "ms_per_field_pass": float(dt.item()), "tiles_bit_identical_to_single_gpu": bool(flag.item()), "n_gpus": world,
            "field_share_per_rank": round(share, 4), "local_region": [R1 - R0, C1 - C0],
            "mse_tiled": out["mse"], "mse_single": mse_single,
            "iterative": {"steps": it.nb_of_deblended_galaxies, "ms_total": float(it_ms.item()), "mse": it.mse},
        }), flush=True)
    assert bool(flag.item()), "tiled residual differs from the single-GPU residual"
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
