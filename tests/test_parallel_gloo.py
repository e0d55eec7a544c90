"""Multi-process (gloo, world_size 2) tests of the sharding logic in debvader_b200.parallel — no GPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from debvader_b200 import parallel as par
from oracle import field_numpy as fo


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # ---- stamp sharding keeps the order contract ---------------------------------------
        imgs = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3)
        full, span = par.deblend_sharded(lambda x: x * 2 + 1, imgs, gather=True)
        assert span == (0, 7) and torch.equal(full, imgs * 2 + 1)
        local, (s, e) = par.deblend_sharded(lambda x: x * 2 + 1, imgs, gather=False)
        assert (s, e) == par.shard_bounds(7, world)[rank] and torch.equal(local, imgs[s:e] * 2 + 1)

        # ---- field tiling + halo exchange of overlapping stamps ----------------------------------
        F, S, C, N = 64, 9, 2, 40
        rng = np.random.default_rng(0)
        field = rng.normal(size=(1, F, F, C))
        cx = rng.integers(-F // 2 + 1, F // 2 - 1, size=N)
        cy = rng.integers(-F // 2 + 1, F // 2 - 1, size=N)
        stamps = rng.random((N, S, S, C)).astype(np.float32)
        off = fo.subtract_offset(F, S)
        x0, y0 = off + cx, off + cy
        owner = par.assign_owners(cx + F // 2, cy + F // 2, F, world)
        touches = par.overlap_matrix(x0, y0, S, F, world)
        mine = np.nonzero(owner == rank)[0]
        got, ids = par.exchange_halo_stamps(torch.from_numpy(stamps[mine]), mine, owner, touches)
        want = np.nonzero(touches[:, rank])[0]
        assert list(ids) == list(want), (rank, ids, want)
        assert np.array_equal(got.numpy(), stamps[want])
        # every rank assembles its own tile from the exchanged stamps, in ascending global index:
        r0, r1, c0, c1 = par.tile_bounds(F, world)[rank]
        full_res = fo.residual_field(field, stamps, cx, cy, cutout_size=S)
        tile = field[0, r0:r1, c0:c1].copy()
        for k, st in zip(ids, got.numpy()):
            canvas = np.zeros((F, F, C))
            fo._paste(canvas, st, int(x0[k]), int(y0[k]), +1)
            tile -= canvas[r0:r1, c0:c1]
        assert np.array_equal(tile, full_res[0, r0:r1, c0:c1]), "tile assembly must be bit-identical to the sequential result"
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_shard_bounds_and_tiles():
    assert par.shard_bounds(10, 4) == [(0, 3), (3, 6), (6, 8), (8, 10)]
    assert par.shard_bounds(0, 2) == [(0, 0), (0, 0)]
    assert par.tile_grid(8) == (2, 4) and par.tile_grid(4) == (2, 2) and par.tile_grid(2) == (1, 2) and par.tile_grid(1) == (1, 1)
    tb = par.tile_bounds(4096, 8)
    assert len(tb) == 8 and tb[0] == (0, 2048, 0, 1024) and tb[-1] == (2048, 4096, 3072, 4096)
    cover = np.zeros((4096, 4096), dtype=np.int8)
    for r0, r1, c0, c1 in tb:
        cover[r0:r1, c0:c1] += 1
    assert (cover == 1).all()
