#!/usr/bin/env python
"""Detection on a field tiled over N GPUs (debvader_b200.detect.detection.TiledDeviceDetector): every rank holds ONLY its owner tile
+ 30-px halo, the ranks exchange the 64x64-mesh statistics (one all-reduce of (2, ny, nx) floats) and their own objects (one
all-gather), and every rank ends with the single-GPU list of centres — checked bit for bit against DeviceDetector on the whole
field (rank 0), then the iterative loop of BASELINE cfg 4 with that detector (IterativeDeblendField(tiled=True, detector="device")).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/detect_tiled_nccl.py [F] [sources]
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
from debvader_b200.deblend_iterative.iterative_deblender import IterativeDeblendField  # noqa: E402
from debvader_b200.detect.detection import DeviceDetector, TiledDeviceDetector  # noqa: E402
from debvader_b200.model.model import load_deblender  # noqa: E402


def blob_field(F, n, seed=6, big=False):
    rng = np.random.default_rng(seed)
    field = rng.standard_normal((1, F, F, 6), dtype=np.float32).astype(np.float64) * 0.03
    yy, xx = np.mgrid[-15:16, -15:16]
    for (px, py) in rng.uniform(40, F - 40, (n, 2)):
        ix, iy = int(px), int(py)
        blob = rng.uniform(0.5, 3.0) * np.exp(-((xx - (px - ix)) ** 2 + (yy - (py - iy)) ** 2) / (2 * rng.uniform(1.2, 2.5) ** 2))
        field[0, iy - 15 : iy + 16, ix - 15 : ix + 16, :] += blob[..., None]
    if big:  # a footprint wider than the halo across the first tile edge: forces the assembled-field path
        Y, X = np.mgrid[0:F, 0:F]
        field[0] += (40.0 * np.exp(-((X - F // 2 - 3) ** 2 + (Y - F // 3) ** 2) / (2 * 14.0 ** 2)))[..., None]
    return field


def main():
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
    rank, world, lrank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lrank)
    dev = torch.device("cuda", lrank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    out = {"n_gpus": world, "field": f"{F}x{F}x6 f64", "sources": N}
    for tag, big in (("plain", False), ("wide_object", True)):
        field = blob_field(F, N, big=big)
        local = par.LocalField.from_full(field, rank, world, device=dev)
        det = TiledDeviceDetector(device=dev)
        c = det(local.data, local)
        c = det(local.data, local)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            c = det(local.data, local)
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / 3 * 1e3], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        # the single-GPU list (rank 0, whole field), broadcast for the comparison
        if rank == 0:
            single = DeviceDetector(device=dev)
            full = torch.from_numpy(field).to(dev)
            ref = single(full)
            del single, full
            shape = torch.tensor([len(ref)], device=dev)
        else:
            shape = torch.zeros(1, dtype=torch.int64, device=dev)
        if world > 1:
            dist.broadcast(shape, 0)
        ref_t = torch.from_numpy(ref).to(dev) if rank == 0 else torch.empty((int(shape.item()), 2), dtype=torch.float64, device=dev)
        if world > 1:
            dist.broadcast(ref_t, 0)
        same = torch.tensor([int(c.shape == tuple(ref_t.shape) and np.array_equal(c, ref_t.cpu().numpy()))], device=dev)
        if world > 1:
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
        out[tag] = {"ms_per_detection_max_over_ranks": float(t.item()), "objects": int(len(c)), "identical_to_single_gpu_on_every_rank": bool(same.item()),
                    "assembled_field_fallbacks": det.fallbacks, "region_share": round(local.data.numel() / (F * F * 6), 4)}
        torch.cuda.empty_cache()
    # the iterative loop on the tiled field with the tiled detector (bounded: random-init weights do not converge)
    field = blob_field(F, N)
    net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234", seed=0)
    net.sample = False
    inner = TiledDeviceDetector(device=dev)

    class Bounded:
        accepts_local = True

        def __init__(self):
            self.calls = 0

        def __call__(self, f, local):
            self.calls += 1
            return inner(f, local) if self.calls <= 4 else np.zeros((0, 2))

    with contextlib.redirect_stdout(io.StringIO()):
        it = IterativeDeblendField(net, field, detector=Bounded(), tiled=True)
        it.iterative_deblending()
        it = IterativeDeblendField(net, field, detector=Bounded(), tiled=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        it.iterative_deblending()
        torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) * 1e3], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["iterative_tiled_device_detector"] = {"ms_total": float(t.item()), "galaxies_per_step": [int(v) for v in it.nb_of_deblended_galaxies],
                                              "assembled_field_fallbacks": inner.fallbacks}
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
