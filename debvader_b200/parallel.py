"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

Stamps are independent, so ``deblend`` shards by contiguous slices with the weights replicated and
NO collective on the data path; results stay on the owning rank unless the caller asks for a
gather.

Fields are split into owner tiles.  A rank holds ONLY its *local region*: its owner
tile plus a 30-pixel halo (stamp half-width 29, +1 for the one-pixel offset between the extraction
and the subtraction window of an even-sized field), clipped to the field — 1/world of the field
plus the halo, never the whole field.  A source belongs to the tile that contains its centre, so
its 59x59 extraction window lies inside the owner's local region.  At field assembly the predicted
stamps whose window reaches into another rank's region are exchanged — ONE ``all_to_all_single``
over NVLink/NCCL, the only communication — and every rank subtracts, in ascending global source
index, every stamp touching its region.  The regions are therefore bit-identical to the
corresponding part of the single-GPU (and the reference's sequential) residual, halo included: the
next iteration of the iterative loop extracts from them without any further exchange.  The field
MSE is the all-reduced sum of the owner tiles' partial sums.

The reference has no distributed code at all; this module is new.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

HALO = 30  # int(59/2) + 1: see the module docstring


def shard_bounds(n: int, world: int):
    """Contiguous, order-preserving split of range(n): [(start, stop)] per rank (sizes differ by <= 1)."""
    base, rem = divmod(int(n), int(world))
    out, s = [], 0
    for r in range(world):
        e = s + base + (1 if r < rem else 0)
        out.append((s, e))
        s = e
    return out


def _world(group=None):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def deblend_sharded(net_fn, images, group=None, gather=False):
    """Run ``net_fn(images[start:stop])`` on this rank's slice.

    net_fn maps an (n,59,59,6) array/tensor to a tensor (n, ...).  Returns ``(local_result,
    (start, stop))`` or, with gather=True, the full result on every rank in the original order
    (all_gather of padded shards; the padding never reaches the caller).

    Latent sampling: a ``Deblender`` draws the noise of local stamp i of call k from (seed, k, i); ranks that
    share a seed would share their noise.  Give every rank its own seed (``load_deblender(..., seed=rank)``,
    as bench.py does) or call ``seed_for_rank(net)`` once."""
    rank, world = _world(group)
    bounds = shard_bounds(len(images), world)
    s, e = bounds[rank]
    local = net_fn(images[s:e])
    if not gather or world == 1:
        return local, (s, e)
    local = local if isinstance(local, torch.Tensor) else torch.as_tensor(local)
    mx = max(b[1] - b[0] for b in bounds)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: e - s] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: b[1] - b[0]] for p, b in zip(parts, bounds)], dim=0), (0, len(images))


def seed_for_rank(net, base_seed: int = 0, group=None):
    """Fold the rank into a Deblender's sampling seed so that stamps on different GPUs draw independent noise
    (the reference draws independent noise for every stamp)."""
    rank, world = _world(group)
    net.seed(int(base_seed) * 1_000_003 + rank + 1)
    return net


# ---------------------------------------------------------------------------------------------
# field tiling
# ---------------------------------------------------------------------------------------------
def tile_grid(world: int):
    """rows x cols of owner tiles for `world` ranks (8 -> 2x4, 4 -> 2x2, 2 -> 1x2, 1 -> 1x1)."""
    r = int(np.floor(np.sqrt(world)))
    while world % r:
        r -= 1
    return r, world // r


def tile_bounds(field_size: int, world: int):
    """[(r0, r1, c0, c1)] of each rank's owner tile."""
    gr, gc = tile_grid(world)
    rb = shard_bounds(field_size, gr)
    cb = shard_bounds(field_size, gc)
    return [(rb[i][0], rb[i][1], cb[j][0], cb[j][1]) for i in range(gr) for j in range(gc)]


def region_bounds(field_size: int, world: int, halo: int = HALO):
    """[(R0, R1, C0, C1)] of each rank's local region: owner tile + halo, clipped to the field."""
    F_ = int(field_size)
    return [(max(r0 - halo, 0), min(r1 + halo, F_), max(c0 - halo, 0), min(c1 + halo, F_)) for r0, r1, c0, c1 in tile_bounds(F_, world)]


def assign_owners(centre_rows, centre_cols, field_size: int, world: int):
    """Owner rank of each source = the tile containing its (clipped) centre pixel."""
    tb = tile_bounds(field_size, world)
    rows = np.clip(np.asarray(centre_rows, dtype=np.int64), 0, field_size - 1)
    cols = np.clip(np.asarray(centre_cols, dtype=np.int64), 0, field_size - 1)
    owner = np.zeros(len(rows), dtype=np.int64)
    for r, (r0, r1, c0, c1) in enumerate(tb):
        owner[(rows >= r0) & (rows < r1) & (cols >= c0) & (cols < c1)] = r
    return owner


def overlap_matrix(x0, y0, S: int, field_size: int, world: int, halo: int = 0):
    """bool (N, world): does the window [x0,x0+S) x [y0,y0+S) of stamp k touch rank r's tile (halo=0) or local
    region (halo=HALO)?"""
    tb = region_bounds(field_size, world, halo)
    x0 = np.asarray(x0, dtype=np.int64)
    y0 = np.asarray(y0, dtype=np.int64)
    m = np.zeros((len(x0), world), dtype=bool)
    for r, (r0, r1, c0, c1) in enumerate(tb):
        m[:, r] = (x0 < r1) & (x0 + S > r0) & (y0 < c1) & (y0 + S > c0)
    return m


class ExchangePlan:
    """Everything the halo exchange and the subtraction of one rank need besides the stamps themselves, derived from the
    TilePlan alone and uploaded in two small copies BEFORE the pass enqueues its device work: a pageable host-to-device copy
    is ordered after everything already on the stream, so an upload made between the network and the exchange would stall the
    host until the network has finished."""

    def __init__(self, tp, rank: int, world: int, device, region):
        mine = tp.mine(rank)
        self.mine = mine
        pos = np.full(len(tp.owner), -1, dtype=np.int64)
        pos[mine] = np.arange(len(mine))
        if world == 1:
            send_ids = [mine[tp.touches[mine, 0]]]
            recv_ids = list(send_ids)
        else:
            send_ids = [mine[tp.touches[mine, dst]] for dst in range(world)]
            recv_ids = [np.nonzero((tp.owner == src) & tp.touches[:, rank])[0].astype(np.int64) for src in range(world)]
        self.send_counts = [len(i) for i in send_ids]
        self.recv_counts = [len(i) for i in recv_ids]
        ids = np.concatenate(recv_ids) if recv_ids else np.zeros(0, dtype=np.int64)
        order = np.argsort(ids, kind="stable")
        self.ids = ids[order]  # global (accepted-source) indices of the stamps this rank applies, ascending
        self.reorder = not np.array_equal(order, np.arange(len(order)))
        sel = np.concatenate([pos[i] for i in send_ids]) if send_ids else np.zeros(0, dtype=np.int64)
        self.send_all_in_order = world == 1 and len(sel) == len(mine)
        R0, _, C0, _ = region
        idx64 = torch.from_numpy(np.concatenate([sel, order]).astype(np.int64)).to(device)
        xy = torch.from_numpy(np.stack([tp.x0[self.ids] - R0, tp.y0[self.ids] - C0]).astype(np.int32)).to(device)
        self.sel = idx64[: len(sel)]
        self.order = idx64[len(sel) :]
        self.x0, self.y0 = xy[0], xy[1]


def exchange_halo_stamps(local_stamps, local_ids, owner, touches, group=None, plan=None):
    """Halo exchange of overlapping stamps.

    local_stamps (n_local,S,S,C) tensor of the stamps this rank owns, local_ids their global
    indices (ascending).  `owner` (N,) and `touches` (N,world) are known to every rank (they only
    depend on the centres).  Returns (stamps, ids): every stamp whose window touches this rank's
    tile / region, sorted by global index.  One all_to_all_single; no other communication.
    `plan` (ExchangePlan): the index tensors prepared — and uploaded — ahead of the device work."""
    rank, world = _world(group)
    if plan is not None:
        if world == 1:
            return (local_stamps if plan.send_all_in_order else local_stamps.index_select(0, plan.sel)), plan.ids
        per = int(np.prod(local_stamps.shape[1:]))
        send = local_stamps.index_select(0, plan.sel).reshape(-1) if len(plan.sel) else local_stamps.new_zeros((0,))
        recv = local_stamps.new_empty((sum(plan.recv_counts) * per,))
        dist.all_to_all_single(recv, send, output_split_sizes=[n * per for n in plan.recv_counts],
                               input_split_sizes=[n * per for n in plan.send_counts], group=group)
        stamps = recv.reshape((-1,) + tuple(local_stamps.shape[1:]))
        if plan.reorder:
            stamps = stamps.index_select(0, plan.order)
        return stamps, plan.ids
    local_ids = np.asarray(local_ids, dtype=np.int64)
    dev = local_stamps.device
    if world == 1:
        keep = touches[local_ids, 0]
        return local_stamps[torch.as_tensor(np.nonzero(keep)[0], device=dev)], local_ids[keep]
    pos = np.full(len(owner), -1, dtype=np.int64)
    pos[local_ids] = np.arange(len(local_ids))
    send_ids = [local_ids[touches[local_ids, dst]] for dst in range(world)]
    recv_ids = [np.nonzero((owner == src) & touches[:, rank])[0].astype(np.int64) for src in range(world)]
    per = int(np.prod(local_stamps.shape[1:]))
    n_send = sum(len(i) for i in send_ids)
    if n_send:
        sel = torch.from_numpy(np.concatenate([pos[ids] for ids in send_ids])).to(dev)
        send = local_stamps.index_select(0, sel).reshape(-1)
    else:
        send = local_stamps.new_zeros((0,))
    recv = local_stamps.new_empty((sum(len(i) for i in recv_ids) * per,))
    dist.all_to_all_single(recv, send, output_split_sizes=[len(i) * per for i in recv_ids],
                           input_split_sizes=[len(i) * per for i in send_ids], group=group)
    ids = np.concatenate(recv_ids) if recv_ids else np.zeros(0, dtype=np.int64)
    stamps = recv.reshape((-1,) + tuple(local_stamps.shape[1:]))
    order = np.argsort(ids, kind="stable")
    if not np.array_equal(order, np.arange(len(order))):
        stamps = stamps.index_select(0, torch.from_numpy(order).to(dev))
    return stamps, ids[order]


class LocalField:
    """A rank's share of a (1,F,F,C) field: its owner tile plus a HALO-pixel border (clipped to the field).

    ``data`` is the (1, R1-R0, C1-C0, C) device tensor — the only part of the field this rank ever holds on its GPU."""

    def __init__(self, data, field_size: int, rank: int, world: int, halo: int = HALO):
        self.field_size, self.rank, self.world, self.halo = int(field_size), int(rank), int(world), int(halo)
        self.tile = tile_bounds(self.field_size, self.world)[self.rank]
        self.region = region_bounds(self.field_size, self.world, self.halo)[self.rank]
        R0, R1, C0, C1 = self.region
        if tuple(data.shape[:3]) != (1, R1 - R0, C1 - C0):
            raise ValueError(f"local region of rank {rank}/{world} must have shape (1,{R1 - R0},{C1 - C0},C), got {tuple(data.shape)}")
        self.data = data

    @classmethod
    def from_full(cls, field_image, rank: int, world: int, device=None, halo: int = HALO):
        """Cut this rank's region out of a full (1,F,F,C) field (host ndarray / memmap or tensor): only the region
        is uploaded."""
        from . import _fieldops

        F_ = int(field_image.shape[1])
        R0, R1, C0, C1 = region_bounds(F_, world, halo)[rank]
        part = field_image[:, R0:R1, C0:C1, :]
        if isinstance(part, torch.Tensor):
            part = part.contiguous() if not part.is_contiguous() else part.clone()
        return cls(_fieldops.to_device_field(part, device), F_, rank, world, halo)

    def like(self, data):
        return LocalField(data, self.field_size, self.rank, self.world, self.halo)

    @property
    def origin(self):
        return self.region[0], self.region[2]

    def owner_slice(self):
        """(row slice, col slice) of the owner tile inside the local region."""
        r0, r1, c0, c1 = self.tile
        R0, _, C0, _ = self.region
        return slice(r0 - R0, r1 - R0), slice(c0 - C0, c1 - C0)

    def owner_tile(self, data=None):
        rs, cs = self.owner_slice()
        return (self.data if data is None else data)[0, rs, cs]

    def nbytes(self):
        return self.data.numel() * self.data.element_size()


class TilePlan:
    """Everything about one tiled pass that depends only on the centres — identical on every rank, computed on the host:
    accepted sources (the order contract: ascending index = detection order), their extraction / subtraction windows,
    owners and the overlap matrix."""

    def __init__(self, galaxy_distances_to_center, field_size: int, world: int, cutout_size: int = 59, halo: int = HALO,
                 touch_halo: bool = True):
        from . import _fieldops

        F_, S = int(field_size), int(cutout_size)
        self.field_size, self.S, self.world = F_, S, world
        self.plan = _fieldops.plan_windows(galaxy_distances_to_center, S, F_)
        ok = self.plan["ok"]
        self.n_sources = len(ok)
        self.idx = np.nonzero(ok)[0]  # accepted sources, in the order the centres were given
        sx, sy, lx, ly = (self.plan[k][self.idx] for k in ("sx", "sy", "lx", "ly"))
        if ((lx != S) | (ly != S)).any():
            raise NotImplementedError("tiled fields take plain in-bounds windows only (a broadcast length-1 window was accepted by the planner)")
        half = int(S / 2)
        gdc = galaxy_distances_to_center
        if isinstance(gdc, np.ndarray) and gdc.ndim == 2 and gdc.dtype != object:  # the usual case: no per-source Python work
            centres = np.asarray(gdc[:, :2], dtype=np.float64)
        else:
            centres = np.asarray([np.asarray(c, dtype=np.float64)[:2] for c in gdc], dtype=np.float64).reshape(-1, 2)
        want_x = -half + np.trunc(centres[self.idx, 0]).astype(np.int64) + int(F_ / 2)
        want_y = -half + np.trunc(centres[self.idx, 1]).astype(np.int64) + int(F_ / 2)
        if not (np.array_equal(want_x, sx) and np.array_equal(want_y, sy)):
            raise NotImplementedError("tiled fields take plain in-bounds windows only (a negative-index wrap-around window was accepted by the planner)")
        self.sx, self.sy = sx, sy
        off = _fieldops.subtract_offset(F_, S)
        self.x0 = off + _fieldops.integer_positions(centres[self.idx, 0], np.zeros(len(self.idx)), "x positions")
        self.y0 = off + _fieldops.integer_positions(centres[self.idx, 1], np.zeros(len(self.idx)), "y positions")
        self.owner = assign_owners(sx + half, sy + half, F_, world)
        self.touches = overlap_matrix(self.x0, self.y0, S, F_, world, halo if touch_halo else 0)

    def mine(self, rank: int):
        """positions (into idx) of the sources rank owns, ascending."""
        return np.nonzero(self.owner == rank)[0]


def extract_local(local: LocalField, tp: TilePlan, mine, nb_of_bands: int, out_dtype=torch.float32):
    """Gather this rank's sources from its local region (extract/extraction.py:21-36 with region-relative windows)."""
    from . import _fieldops

    R0, R1, C0, C1 = local.region
    sx, sy = tp.sx[mine] - R0, tp.sy[mine] - C0
    S = tp.S
    if len(mine) and (sx.min() < 0 or sy.min() < 0 or (sx + S).max() > R1 - R0 or (sy + S).max() > C1 - C0):
        raise RuntimeError("an owned source's window leaves the local region (halo too small)")
    n = len(mine)
    sub = {"sx": sx, "sy": sy, "lx": np.full(n, S), "ly": np.full(n, S), "ok": np.ones(n, dtype=bool)}
    cut, _ = _fieldops.extract(local.data, sub, S, nb_of_bands, out_dtype=out_dtype)
    return cut


def subtract_local(local: LocalField, tp: TilePlan, stamps, ids, alpha: float = -1.0, out=None, base="field", plan=None):
    """region + alpha * (every exchanged stamp, ascending global index), clipped to the region
    (deblend/field_deblender.py:46-97 restricted to the local region).  base="zeros" starts from zeros (predicted fields).
    `plan` (ExchangePlan): region-relative window positions already on the device."""
    from . import _fieldops

    R0, _, C0, _ = local.region
    src = local.data if base == "field" else None
    x0, y0 = (plan.x0, plan.y0) if plan is not None else (tp.x0[ids] - R0, tp.y0[ids] - C0)
    return _fieldops.window_axpy(src, stamps.contiguous(), x0, y0, alpha, out=out,
                                 field_shape=tuple(local.data.shape), dtype=local.data.dtype)


def field_mse_tiled(local: LocalField, a, b, group=None) -> float:
    """training/metrics.py:4-12 over the WHOLE field from the owner tiles: local partial sum + one all-reduce."""
    from . import _fieldops

    (rs, cs) = local.owner_slice()
    part = _fieldops.sqdiff_sum_rect(a, b, rs.start, rs.stop, cs.start, cs.stop)
    if local.world > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    F_ = local.field_size
    return float(part.item()) / float(F_ * F_ * a.shape[-1])


def gather_field(local: LocalField, data=None, group=None):
    """Assemble the full (1,F,F,C) field on every rank's HOST from the owner tiles (for callers that want the
    reference's ndarray; the device never holds more than the local region plus one padded tile per rank)."""
    data = local.data if data is None else data
    tile = local.owner_tile(data).contiguous()
    F_, Cc = local.field_size, data.shape[-1]
    tb = tile_bounds(F_, local.world)
    if local.world == 1:
        return tile.cpu().numpy()[None]
    mh = max(b[1] - b[0] for b in tb)
    mw = max(b[3] - b[2] for b in tb)
    pad = torch.zeros((mh, mw, Cc), dtype=tile.dtype, device=tile.device)
    pad[: tile.shape[0], : tile.shape[1]] = tile
    parts = [torch.empty_like(pad) for _ in range(local.world)]
    dist.all_gather(parts, pad, group=group)
    out = np.empty((1, F_, F_, Cc), dtype=tile.cpu().numpy().dtype)
    for p, (r0, r1, c0, c1) in zip(parts, tb):
        out[0, r0:r1, c0:c1] = p[: r1 - r0, : c1 - c0].cpu().numpy()
    return out


def deblend_field_tiled(net, field_image, galaxy_distances_to_center, group=None, cutout_size=59, nb_of_bands=6, sample=False,
                        seed=None, device=None):
    """One deblending pass over a field split into owner tiles, one rank per GPU (BASELINE config 4).

    ``field_image`` is the full (1,F,F,C) field on the host (only this rank's local region is uploaded) or an
    already cut ``LocalField``.  The rank deblends the sources whose centre lies in ITS tile (extract -> net), the
    predicted stamps reaching into other regions are exchanged (``exchange_halo_stamps``) and the rank subtracts every
    stamp that touches its region in ascending global source index.  Returns ``(residual LocalField, accepted source
    indices, TilePlan)``; the regions of all ranks are bit-identical to the single-GPU residual field.
    ``sample=False`` (z = loc) makes the pass deterministic."""
    rank, world = _world(group)
    local = field_image if isinstance(field_image, LocalField) else LocalField.from_full(field_image, rank, world, device)
    tp = TilePlan(galaxy_distances_to_center, local.field_size, world, cutout_size)
    xp = ExchangePlan(tp, rank, world, local.data.device, local.region)
    mine = xp.mine
    cut = extract_local(local, tp, mine, nb_of_bands, out_dtype=torch.float32)
    if len(mine):
        mean = net(cut, sample=sample, seed=seed).mean().tensor
    else:
        mean = torch.empty((0, tp.S, tp.S, nb_of_bands), device=local.data.device, dtype=torch.float32)
    stamps, ids = exchange_halo_stamps(mean.contiguous(), mine, tp.owner, tp.touches, group, plan=xp)
    res = subtract_local(local, tp, stamps, ids, -1.0, plan=xp) if len(ids) else local.data.clone()
    return local.like(res), [int(i) for i in tp.idx], tp
