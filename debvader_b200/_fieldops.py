"""Host side of the field operators: index planning (bit-identical to the reference by
construction) and thin wrappers that enqueue the CUDA kernels through the C-ABI.

Index planning restates extract/extraction.py:26-32 literally (int() truncation toward zero,
Python slice clipping / negative wrap-around via the same arithmetic as ``slice.indices``, numpy's
"length S or length 1 broadcasts" assignment rule), vectorised for ndarray input.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _ffi

_DT = {torch.float32: _ffi.F32, torch.float64: _ffi.F64}


# ---------------------------------------------------------------------------------------------
# planning (host, integer-exact)
# ---------------------------------------------------------------------------------------------
def _slice_norm(v, n):
    """start/stop normalisation of slice(start, stop).indices(n) for step 1."""
    v = np.where(v < 0, v + n, v)
    return np.clip(v, 0, n)


def plan_windows(galaxy_distances_to_center, cutout_size: int, field_size: int):
    """For every centre: source window start (row, col), its lengths, and acceptance.

    Returns dict of arrays: sx, sy (int64 start row/col), lx, ly (int64), ok (bool).
    """
    S, F_ = int(cutout_size), int(field_size)
    half = int(S / 2)
    fh = int(F_ / 2)
    n = len(galaxy_distances_to_center)
    xi = np.zeros(n, dtype=np.int64)
    yi = np.zeros(n, dtype=np.int64)
    valid = np.ones(n, dtype=bool)
    arr = None
    if isinstance(galaxy_distances_to_center, np.ndarray) and galaxy_distances_to_center.dtype.kind in "iuf":
        arr = galaxy_distances_to_center
    if arr is not None and arr.ndim == 2 and arr.shape[1] >= 2 and (arr.dtype.kind in "iu" or np.isfinite(arr[:, :2]).all()):
        xi = np.trunc(arr[:, 0]).astype(np.int64)  # int() truncates toward zero
        yi = np.trunc(arr[:, 1]).astype(np.int64)
    else:
        for i in range(n):
            try:
                c = galaxy_distances_to_center[i]
                xi[i], yi[i] = int(c[0]), int(c[1])
            except ValueError:  # int(nan): numpy raises ValueError, which the reference swallows (extraction.py:35)
                valid[i] = False
    xs = -half + xi + fh
    ys = -half + yi + fh
    W = 2 * half + 1  # x_end - x_start (extraction.py:27)
    sx, ex = _slice_norm(xs, F_), _slice_norm(xs + W, F_)
    sy, ey = _slice_norm(ys, F_), _slice_norm(ys + W, F_)
    lx = np.maximum(ex - sx, 0)
    ly = np.maximum(ey - sy, 0)
    ok = valid & ((lx == S) | (lx == 1)) & ((ly == S) | (ly == 1))
    return {"sx": sx, "sy": sy, "lx": lx, "ly": ly, "ok": ok}


def subtract_offset(field_size: int, cutout_size: int) -> int:
    """pos_offset of get_residual_field (deblend/field_deblender.py:72)."""
    return int((int(field_size) - int(cutout_size)) / 2)


def positions(dist, shifts):
    """x_pos = distance + shift (field_deblender.py:83-90) as float64, and whether all are integer-valued."""
    p = np.asarray(dist, dtype=np.float64) + np.asarray(shifts, dtype=np.float64)
    if not np.isfinite(p).all():
        raise ValueError("positions must be finite")
    return p, bool(np.array_equal(np.rint(p), p))


def integer_positions(dist, shifts, what="positions"):
    """positions() as int64 for the callers that only take whole pixels (the tiled multi-GPU pass)."""
    p, integer = positions(dist, shifts)
    if not integer:
        raise NotImplementedError(f"{what} are not integer-valued: this path places stamps on whole pixels only")
    return np.rint(p).astype(np.int64)


SPLINE_MARGIN = 28  # pixels kept around the stamp by the sub-pixel placement (prefilter response 0.268^28 = 1e-16)


# ---------------------------------------------------------------------------------------------
# kernels
# ---------------------------------------------------------------------------------------------
def _require_cuda(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda):
        raise TypeError(f"{name} must be a CUDA tensor")
    if t.dtype not in _DT:
        raise TypeError(f"{name} must be float32 or float64, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name} must be contiguous")


def to_device_field(field_image, device=None):
    """(1,F,F,C) array-like -> contiguous CUDA tensor, dtype kept (float64 for the reference's fields)."""
    if isinstance(field_image, torch.Tensor):
        t = field_image
    else:
        a = np.asarray(field_image)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float64)
        t = torch.from_numpy(np.ascontiguousarray(a))
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device())
    return t.to(device).contiguous()


def extract(field_dev, plan, cutout_size: int, nb_of_bands: int, out_dtype=torch.float64):
    """Gather the accepted windows of `plan` into a zero-initialised (N,S,S,nb) CUDA tensor."""
    _require_cuda(field_dev, "field")
    S = int(cutout_size)
    n = len(plan["ok"])
    F_ = field_dev.shape[2]  # row pitch in pixels (the field may be a rectangular local region: the planner owns bounds)
    Cf = field_dev.shape[3]
    ok = plan["ok"]
    dev = field_dev.device
    shape = (n, S, S, nb_of_bands)
    if Cf != nb_of_bands:
        if Cf != 1:  # numpy cannot broadcast (.., Cf) into (.., nb): every stamp raises ValueError in the reference
            return torch.zeros(shape, device=dev, dtype=out_dtype), []
        field_dev = field_dev.expand(1, field_dev.shape[1], F_, nb_of_bands).contiguous()
    # accepted indices (and their Python-list form, the reference's list_idx) are cached on the plan: a repeated extraction
    # is then one kernel launch, not an O(n) walk on the host
    if "_idx" not in plan:
        plan["_idx"] = np.nonzero(ok)[0]
        plan["_idx_list"] = [int(i) for i in plan["_idx"]]
    idx = plan["_idx"]
    # rejected stamps stay zero like the reference's np.zeros; when all are accepted skip the memset
    out = torch.empty(shape, device=dev, dtype=out_dtype) if idx.size == n else torch.zeros(shape, device=dev, dtype=out_dtype)
    if idx.size == 0:
        return out, []
    # one packed upload of the plan (cached on the plan: repeated extractions reuse the device copy)
    cache = plan.get("_dev")
    if cache is None or cache[0] != str(dev):
        m = idx.size
        flags = ((plan["lx"][idx] == 1) & (S != 1)).astype(np.uint8) | (((plan["ly"][idx] == 1) & (S != 1)).astype(np.uint8) << 1)
        buf = np.zeros(4 * m + (m + 3) // 4, dtype=np.int32)
        buf[:m] = plan["sx"][idx]
        buf[m : 2 * m] = plan["sy"][idx]
        buf[2 * m : 4 * m] = idx.astype(np.int64).view(np.int32)
        buf[4 * m :].view(np.uint8)[:m] = flags
        t = torch.from_numpy(buf).to(dev)
        cache = (str(dev), t[:m], t[m : 2 * m], t[4 * m :].view(torch.uint8), t[2 * m : 4 * m].view(torch.int64))
        plan["_dev"] = cache
    _, sx, sy, fl, slot = cache
    with torch.cuda.device(dev):
        _ffi.check(
            _ffi.lib().dbv_extract(_ffi.ptr(field_dev), _DT[field_dev.dtype], F_, nb_of_bands, _ffi.ptr(sx), _ffi.ptr(sy), _ffi.ptr(fl),
                                   _ffi.ptr(slot), int(idx.size), S, _ffi.ptr(out), _DT[out_dtype], _ffi.stream_ptr())
        )
    return out, list(plan["_idx_list"])


def window_axpy(field_in, stamps, x0, y0, alpha: float, out=None, field_shape=None, dtype=torch.float64, planar=False):
    """out = field_in + alpha * sum_k paste(stamps[k] at (x0[k], y0[k])), deterministic (see dbv_window_axpy_rect).
    stamps (N,S,S,C), or (N,C,S,S) with planar=True (the windows of spline_place).  The field may be rectangular
    ((1,FH,FW,C): a rank's local region of a tiled field; positions are relative to it and windows are clipped).
    ``out is field_in`` is the in-place form: only the covered elements are read and written."""
    _require_cuda(stamps, "stamps")
    dev = stamps.device
    if field_in is not None:
        _require_cuda(field_in, "field")
        shape, dtype = tuple(field_in.shape), field_in.dtype
    else:
        shape = tuple(field_shape)
    FH, FW, Cc = shape[-3], shape[-2], shape[-1]
    n, S = stamps.shape[0], stamps.shape[2 if planar else 1]
    if out is None:
        out = torch.empty(shape, device=dev, dtype=dtype)
    else:
        _require_cuda(out, "out")
        if tuple(out.shape) != shape or out.dtype != dtype:
            raise ValueError("out must have the field's shape and dtype")
    if n == 0:
        if field_in is None:
            out.zero_()
        elif out is not field_in:
            out.copy_(field_in)
        return out
    if isinstance(x0, torch.Tensor) and isinstance(y0, torch.Tensor):  # positions already on the device (int32)
        if not (x0.is_cuda and y0.is_cuda and x0.dtype == torch.int32 and y0.dtype == torch.int32 and x0.numel() == n == y0.numel()):
            raise TypeError("device positions must be int32 CUDA tensors of length N")
        xs, ys = x0.contiguous(), y0.contiguous()
    else:  # one packed upload
        xy = torch.from_numpy(np.stack([np.asarray(x0, dtype=np.int32).reshape(-1), np.asarray(y0, dtype=np.int32).reshape(-1)])).to(dev)
        xs, ys = xy[0], xy[1]
    lib = _ffi.lib()
    with torch.cuda.device(dev):
        # binning scratch from torch's stream-ordered caching allocator: private to this call's stream
        sb = int(lib.dbv_window_axpy_scratch_bytes(FH, FW))
        scratch = torch.empty((sb // 4,), device=dev, dtype=torch.int32)
        _ffi.check(
            lib.dbv_window_axpy_rect(_ffi.ptr(field_in), _ffi.ptr(out), _DT[dtype], FH, FW, Cc, _ffi.ptr(stamps), _DT[stamps.dtype],
                                     int(bool(planar)), _ffi.ptr(xs), _ffi.ptr(ys), n, S, float(alpha), _ffi.ptr(scratch), sb,
                                     _ffi.stream_ptr())
        )
    return out


def spline_extent(size: int, margin: int = SPLINE_MARGIN) -> int:
    """side of the window dbv_spline_place writes for a data block of `size` samples per axis."""
    try:
        return _ffi.check(_ffi.lib().dbv_spline_extent(int(size), int(margin)))
    except _ffi.DbvError as e:
        raise NotImplementedError(f"sub-pixel placement: {e}") from None


def _anchor(origin, pos, margin):
    lim = 2**30
    return (np.asarray(origin, dtype=np.int64) - margin - 1 + np.clip(np.floor(pos), -lim, lim).astype(np.int64)).astype(np.int32)


def spline_place(data, pos_x, pos_y, field_size: int, margin: int = SPLINE_MARGIN, origin_x=None, origin_y=None):
    """scipy.ndimage.shift of a zero canvas holding `data` (field_deblender.py:66-95), on the window that matters.

    data (N,S,S,C) CUDA f32/f64, sample 0 at canvas (origin_x[k], origin_y[k]) — default int((F-S)/2), the padded
    stamp of the reference; pos = the shift (float).  Returns (placed (N,C,E,E) f64 — planar —, ax, ay): window k covers
    field rows ax[k]..ax[k]+E, cols ay[k]..ay[k]+E.
    """
    _require_cuda(data, "data")
    dev = data.device
    n, S, _, Cc = data.shape
    E = spline_extent(S, margin)
    px = np.ascontiguousarray(pos_x, dtype=np.float64)
    py = np.ascontiguousarray(pos_y, dtype=np.float64)
    off = subtract_offset(field_size, S)
    ox = np.full(n, off, dtype=np.int64) if origin_x is None else np.asarray(origin_x, dtype=np.int64)
    oy = np.full(n, off, dtype=np.int64) if origin_y is None else np.asarray(origin_y, dtype=np.int64)
    ax, ay = _anchor(ox, px, margin), _anchor(oy, py, margin)
    pos = torch.from_numpy(np.stack([px, py])).to(dev)
    ints = torch.from_numpy(np.stack([ax, ay, ox.astype(np.int32), oy.astype(np.int32)])).to(dev)
    scratch = torch.empty((int(_ffi.lib().dbv_spline_scratch_doubles(n, S, Cc, int(margin))),), device=dev, dtype=torch.float64)
    placed = torch.empty((n, Cc, E, E), device=dev, dtype=torch.float64)
    with torch.cuda.device(dev):
        _ffi.check(
            _ffi.lib().dbv_spline_place(_ffi.ptr(data), _DT[data.dtype], n, S, Cc, int(field_size), int(off),
                                        None if origin_x is None else _ffi.ptr(ints[2]), None if origin_y is None else _ffi.ptr(ints[3]),
                                        _ffi.ptr(pos[0]), _ffi.ptr(pos[1]), _ffi.ptr(ints[0]), _ffi.ptr(ints[1]), int(margin),
                                        _ffi.ptr(scratch), _ffi.ptr(placed), _ffi.stream_ptr())
        )
    return placed, ax, ay


def spline_window_axpy(field_in, stamps, pos_x, pos_y, alpha: float, field_shape=None, dtype=torch.float64, batch: int = 2048,
                       margin: int = SPLINE_MARGIN):
    """out = field_in + alpha * sum_k ndimage.shift(padded stamps[k], (pos_x[k], pos_y[k])), stamps applied in ascending k
    (batches of `batch` stamps bound the memory of the placed windows — 0.66 MB each for DC2 plus 0.34 MB of scratch —; every
    batch is one more pass over the field, so the default covers a whole 2000-source field in one)."""
    _require_cuda(stamps, "stamps")
    if field_in is not None:
        shape, dtype = tuple(field_in.shape), field_in.dtype
    else:
        shape = tuple(field_shape)
    F_ = shape[-3]
    n = stamps.shape[0]
    out = None
    for b0 in range(0, max(n, 1), batch):
        sl = slice(b0, min(b0 + batch, n))
        placed, ax, ay = spline_place(stamps[sl], pos_x[sl], pos_y[sl], F_, margin)
        src = field_in if out is None else out
        out = window_axpy(src, placed, ax, ay, alpha, out=out, field_shape=shape, dtype=dtype, planar=True)
    return out


def center_mse(cutouts, means, lo: int, hi: int):
    """(N,) float64 CUDA tensor of the centre-window MSE (field_deblender.py:323-332)."""
    _require_cuda(cutouts, "cutouts")
    _require_cuda(means, "means")
    n, S, _, Cc = cutouts.shape
    out = torch.empty((n,), device=cutouts.device, dtype=torch.float64)
    with torch.cuda.device(cutouts.device):
        _ffi.check(
            _ffi.lib().dbv_center_mse(_ffi.ptr(cutouts), _DT[cutouts.dtype], _ffi.ptr(means), n, S, Cc, int(lo), int(hi), _ffi.ptr(out),
                                      _ffi.stream_ptr())
        )
    return out


def mse(a, b) -> float:
    """training/metrics.py:4-12 on the device (uploads host arrays)."""
    dev = a.device if isinstance(a, torch.Tensor) and a.is_cuda else (b.device if isinstance(b, torch.Tensor) and b.is_cuda else None)
    ta, tb = to_device_field(a, dev), to_device_field(b, dev)
    if ta.shape != tb.shape:
        ta, tb = torch.broadcast_tensors(ta, tb)
        ta, tb = ta.contiguous(), tb.contiguous()
    if ta.dtype != tb.dtype:
        ta, tb = ta.double(), tb.double()
    n = ta.numel()
    lib = _ffi.lib()
    sb = int(lib.dbv_mse_scratch_bytes())
    scratch = torch.empty((sb // 8,), device=ta.device, dtype=torch.float64)
    out = torch.empty((1,), device=ta.device, dtype=torch.float64)
    with torch.cuda.device(ta.device):
        _ffi.check(lib.dbv_mse(_ffi.ptr(ta), _ffi.ptr(tb), _DT[ta.dtype], n, _ffi.ptr(out), _ffi.ptr(scratch), sb, _ffi.stream_ptr()))
    return float(out.item())


def sqdiff_sum_rect(a, b, r0: int, r1: int, c0: int, c1: int) -> torch.Tensor:
    """sum((a-b)^2) over rows [r0,r1) x cols [c0,c1) (all bands) of two (1,H,W,C) CUDA tensors of the same shape:
    the owner-tile partial sum of a tiled field MSE.  Returns a (1,) float64 CUDA tensor (no host sync)."""
    _require_cuda(a, "a")
    _require_cuda(b, "b")
    if a.shape != b.shape or a.dtype != b.dtype:
        raise ValueError("a and b must have the same shape and dtype")
    H, W, Cc = a.shape[-3], a.shape[-2], a.shape[-1]
    if not (0 <= r0 < r1 <= H and 0 <= c0 < c1 <= W):
        raise ValueError("bad sub-rectangle")
    lib = _ffi.lib()
    sb = int(lib.dbv_mse_scratch_bytes())
    scratch = torch.empty((sb // 8,), device=a.device, dtype=torch.float64)
    out = torch.empty((1,), device=a.device, dtype=torch.float64)
    esz = a.element_size()
    off = (r0 * W + c0) * Cc * esz
    with torch.cuda.device(a.device):
        _ffi.check(lib.dbv_sqdiff_sum_rect(C.c_void_p(a.data_ptr() + off), C.c_void_p(b.data_ptr() + off), _DT[a.dtype], r1 - r0,
                                           (c1 - c0) * Cc, W * Cc, W * Cc, _ffi.ptr(out), _ffi.ptr(scratch), sb, _ffi.stream_ptr()))
    return out
