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
        # the same exchange through a prepared ExchangePlan (index tensors made — on a GPU: uploaded — ahead of the device work)
        import types

        tp = types.SimpleNamespace(owner=owner, touches=touches, x0=x0, y0=y0, mine=lambda r: np.nonzero(owner == r)[0])
        region = par.region_bounds(F, world, 0)[rank]
        xp = par.ExchangePlan(tp, rank, world, torch.device("cpu"), region)
        got2, ids2 = par.exchange_halo_stamps(torch.from_numpy(stamps[mine]), mine, owner, touches, plan=xp)
        assert list(ids2) == list(want) and np.array_equal(got2.numpy(), stamps[want])
        assert np.array_equal(xp.x0.numpy(), (x0[want] - region[0]).astype(np.int32)) and np.array_equal(xp.y0.numpy(), (y0[want] - region[2]).astype(np.int32))
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _tiled_worker(rank, world, port, q):
    """DeblendField(tiled=True) / IterativeDeblendField(tiled=True) under gloo with the oracle-backed CPU stand-ins of the
    device operators: every rank holds its owner tile + halo only and must reproduce the single-process result."""
    import sys

    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import cpu_ops
    from golden.make_golden import fake_net

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        class MP:  # minimal monkeypatch
            def setattr(self, obj, name, val):
                setattr(obj, name, val)

        cpu_ops.install(MP())
        from debvader_b200.deblend.field_deblender import DeblendField
        from debvader_b200.deblend_iterative.iterative_deblender import IterativeDeblendField

        for F in (200, 201):  # even F: subtraction window one pixel up-left of the extraction window
            rng = np.random.default_rng(F)
            field = rng.normal(0, 0.3, (1, F, F, 6))
            centres = rng.integers(-(F // 2) - 5, F // 2 + 5, size=(60, 2)).astype(np.float64)  # some fall off the field
            single = DeblendField(fake_net, field)
            rec1 = single.deblend_field(centres, mse_criterion=0.05)
            res1 = single.get_residual_field()
            pred1 = single.get_predicted_field()

            obj = DeblendField(fake_net, field, tiled=True)
            loc = obj._local
            R0, R1, C0, C1 = loc.region
            assert loc.data.shape == (1, R1 - R0, C1 - C0, 6) and loc.nbytes() < field.nbytes * (1.0 / world + 0.45)
            rec = obj.deblend_field(centres, mse_criterion=0.05)
            # this rank's records = the single-process records of the sources it owns, list_idx global, order kept
            want_rows = [i for i, k in enumerate(rec1["list_idx"]) if k in set(rec["list_idx"])]
            assert list(rec["list_idx"]) == [int(rec1["list_idx"][i]) for i in want_rows]
            assert list(rec["passed_cuts"]) == [bool(rec1["passed_cuts"][i]) for i in want_rows]
            for a, i in zip(rec["output_images_mean"], want_rows):
                assert np.array_equal(np.asarray(a), np.asarray(rec1["output_images_mean"][i]))
            assert obj.nb_of_deblended_galaxies == single.nb_of_deblended_galaxies
            # the whole local region (halo included) is bit-identical to the single-process residual
            reg = obj.get_residual_field(as_tensor=True)
            assert np.array_equal(reg.numpy(), res1[:, R0:R1, C0:C1]), "tiled residual region differs"
            full = obj.get_residual_field()
            assert np.array_equal(full, res1), "gathered tiled residual differs"
            pm = obj.get_predicted_field()["predicted_mean_field"]
            assert np.array_equal(pm, pred1["predicted_mean_field"])
            m = obj.field_mse(obj.field_tensor, reg)
            assert abs(m - np.mean((field - res1) ** 2)) <= 1e-13 * max(m, 1e-30)

        # iterative loop on a tiled field: same control flow, same records as the single-process loop
        F = 160
        rng = np.random.default_rng(5)
        field = rng.normal(0, 0.1, (1, F, F, 6))
        steps = [rng.integers(-40, 40, size=(n, 2)).astype(np.float64) for n in (5, 9, 4)]

        def make_detector():
            calls = []

            def detector(f):
                calls.append(np.asarray(f).copy())
                return steps[min(len(calls) - 1, len(steps) - 1)]

            detector.calls = calls
            return detector

        d1 = make_detector()
        ref = IterativeDeblendField(fake_net, field, detector=d1)
        rec1 = ref.iterative_deblending()
        d2 = make_detector()
        obj = IterativeDeblendField(fake_net, field, detector=d2, tiled=True)
        rec = obj.iterative_deblending()
        assert obj.nb_of_deblended_galaxies == ref.nb_of_deblended_galaxies == [5, 9, 4]
        assert len(d2.calls) == len(d1.calls) == 3
        for a, b in zip(d1.calls, d2.calls):  # the (gathered) field every step's detector saw
            assert np.array_equal(a, b)
        np.testing.assert_allclose(obj.mse, ref.mse, rtol=1e-12)
        mine = [i for i, k in enumerate(rec1["list_idx"]) if k in set(rec["list_idx"])]
        assert list(rec["list_idx"]) == [int(rec1["list_idx"][i]) for i in mine]
        all_idx = [None] * world
        dist.all_gather_object(all_idx, [int(k) for k in rec["list_idx"]])
        assert sorted(sum(all_idx, [])) == [int(k) for k in rec1["list_idx"]]
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, repr(e) + traceback.format_exc()[-1500:]))
    finally:
        dist.destroy_process_group()


def _run(worker, world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res


def test_tiled_field_world2_gloo():
    _run(_tiled_worker, 2)


def test_tiled_field_world4_gloo():
    _run(_tiled_worker, 4)


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
    rb = par.region_bounds(4096, 8)
    assert rb[0] == (0, 2078, 0, 1054) and rb[5] == (2018, 4096, 994, 2078)
    # per-rank share of the field: 1/8 plus the 30-px halo
    assert max((r1 - r0) * (c1 - c0) for r0, r1, c0, c1 in rb) / 4096**2 < 1 / 8 + 0.012
    tb = par.tile_bounds(4096, 8)
    assert len(tb) == 8 and tb[0] == (0, 2048, 0, 1024) and tb[-1] == (2048, 4096, 3072, 4096)
    cover = np.zeros((4096, 4096), dtype=np.int8)
    for r0, r1, c0, c1 in tb:
        cover[r0:r1, c0:c1] += 1
    assert (cover == 1).all()
