"""reference detect/detection.py:5-56 — SExtractor detection on the r band.

The reference delegates to the third-party CPU library `sep` (``sep.Background`` + ``sep.extract``).  Two backends here:

``backend="sep"``     the reference's own call sequence, with a lazy import so that the package works without `sep`
                      (the default for host arrays: the reference's behaviour, library included).
``backend="device"``  the same pipeline as hand-written CUDA kernels on a field that stays on the GPU (SURVEY §8f-3;
                      csrc/detect_kernels.cu through ``dbv_detect``): mesh background, 7x7 matched filter, threshold
                      1.5 x global rms, 8-connected components of >= 4 pixels in the order SExtractor's scan completes them,
                      barycentres -> (row, col) offsets from the field centre.  The default for CUDA tensors, and what
                      ``IterativeDeblendField(..., detector="device")`` uses: no field-sized transfer per iteration.
                      Restated from the published algorithm and bit-exact with oracle/detect_numpy.py; parity with `sep`
                      itself is UNPINNED (not installable where this was built).  Not restated: multi-threshold deblending
                      and sep's `clean` pass — a connected footprint is one detection; the iterative loop finds the other
                      members of a blend in the residual of the next step.
"""
import ctypes as C

import numpy as np

# 7x7 convolution mask of a gaussian PSF with FWHM = 3.0 pixels (detection.py:25-35)
FILTER_KERNEL = np.array(
    [
        [0.004963, 0.021388, 0.051328, 0.068707, 0.051328, 0.021388, 0.004963],
        [0.021388, 0.092163, 0.221178, 0.296069, 0.221178, 0.092163, 0.021388],
        [0.051328, 0.221178, 0.530797, 0.710525, 0.530797, 0.221178, 0.051328],
        [0.068707, 0.296069, 0.710525, 0.951108, 0.710525, 0.296069, 0.068707],
        [0.051328, 0.221178, 0.530797, 0.710525, 0.530797, 0.221178, 0.051328],
        [0.021388, 0.092163, 0.221178, 0.296069, 0.221178, 0.092163, 0.021388],
        [0.004963, 0.021388, 0.051328, 0.068707, 0.051328, 0.021388, 0.004963],
    ]
)
DETECT_THRESH = 1.5  # detection.py:19
MINAREA = 4  # detection.py:22
R_BAND = 2  # detection.py:14


def normalised_taps(kernel=FILTER_KERNEL):
    """float32 taps divided by the float32 running sum of their absolute values (sep normalises the mask it is given)."""
    k = np.ascontiguousarray(kernel, np.float32)
    s = np.float32(0.0)
    for v in k.ravel():
        s = np.float32(s + abs(v))
    return np.ascontiguousarray(k / s, np.float32)


class DeviceDetector:
    """``detector(field) -> (N, 2)`` float64 centres on the GPU.  ``accepts_tensor``: IterativeDeblendField hands it the
    device-resident residual field itself.  Work buffers are kept per field shape; one instance serves one stream at a time."""

    accepts_tensor = True

    def __init__(self, device=None, max_objects=1 << 17, thresh=DETECT_THRESH, minarea=MINAREA, band=R_BAND, kernel=FILTER_KERNEL):
        self.device = device
        self.max_objects = int(max_objects)
        self.thresh, self.minarea, self.band = float(thresh), int(minarea), int(band)
        self.taps = normalised_taps(kernel)
        self._buf = {}
        self.last = None  # details of the last call (device tensors)

    # ---- device plumbing: the five places that touch CUDA.  tests/test_detect_emul.py overrides them to drive ALL of the host logic
    # below (buffers, calls, synchronisation, the tiled exchanges and fall-backs) on CPU tensors with the kernels compiled as host C++.
    def _lib(self):
        from .. import _ffi

        return _ffi.lib()

    def _check(self, rc):
        from .. import _ffi

        return _ffi.check(rc)

    def _device_ctx(self, t):
        import torch

        return torch.cuda.device(t.device)

    def _stream(self):
        from .. import _ffi

        return _ffi.stream_ptr()

    def _require_device(self, t):
        if not t.is_cuda:
            raise ValueError("the device detector needs a CUDA tensor (or a host array to upload)")

    def _buffers(self, H, W, dev):
        import torch

        from .. import _ffi

        key = (H, W, str(dev))
        b = self._buf.get(key)
        if b is None:
            nbytes = int(self._lib().dbv_detect_scratch_bytes(H, W, self.max_objects))
            M = self.max_objects
            b = {
                "scratch": torch.empty(nbytes + 256, dtype=torch.uint8, device=dev),
                "n": torch.zeros(1, dtype=torch.int32, device=dev),
                "xy": torch.empty((M, 2), dtype=torch.float64, device=dev),
                "centres": torch.empty((M, 2), dtype=torch.float64, device=dev),
                "npix": torch.empty(M, dtype=torch.int32, device=dev),
                "stats": torch.zeros(4, dtype=torch.float32, device=dev),
                "nbytes": nbytes,
            }
            off = (-b["scratch"].data_ptr()) % 256
            b["base"] = b["scratch"].data_ptr() + off
            self._buf = {key: b}  # one shape at a time: a 4096^2 field needs ~0.6 GB of planes
        return b

    def run(self, field_image, band=None):
        """Enqueue the detection; returns the buffer dict (device tensors; ``n`` not yet read)."""
        band = self.band if band is None else int(band)
        import torch

        from .. import _ffi

        t = field_image
        if not isinstance(t, torch.Tensor):
            dev = torch.device(self.device if self.device is not None else "cuda")
            t = torch.as_tensor(np.ascontiguousarray(np.asarray(field_image))).to(dev)
        self._require_device(t)
        if t.ndim == 4:
            if t.shape[0] != 1:
                raise ValueError(f"field_image must have shape (1, F, F, C), got {tuple(t.shape)}")
            t = t[0]
        if t.ndim != 3 or t.dtype not in (torch.float64, torch.float32) or not t.is_contiguous():
            raise ValueError("field_image must be a contiguous (1, F, F, C) float64 / float32 tensor")
        H, W, Cn = (int(v) for v in t.shape)
        if not 0 <= band < Cn:
            raise ValueError(f"band {band} outside the field's {Cn} bands")
        b = self._buffers(H, W, t.device)
        with self._device_ctx(t):
            self._check(self._lib().dbv_detect(
                _ffi.ptr(t), 1 if t.dtype == torch.float64 else 0, H, W, W, Cn, band,
                self.taps.ctypes.data_as(C.c_void_p), int(self.taps.shape[0]), int(self.taps.shape[1]), self.thresh, self.minarea,
                int(H / 2), int(W / 2), self.max_objects, C.c_void_p(b["base"]), b["nbytes"], _ffi.ptr(b["n"]), _ffi.ptr(b["xy"]),
                _ffi.ptr(b["centres"]), _ffi.ptr(b["npix"]), _ffi.ptr(b["stats"]), self._stream()))
        b["shape"] = (H, W)
        self.last = b
        return b

    def __call__(self, field_image, return_details=False, band=None):
        b = self.run(field_image, band)
        n = int(b["n"].item())  # the one synchronisation: the host index planner needs the centres anyway
        if n > self.max_objects:
            raise RuntimeError(f"{n} objects detected, more than max_objects={self.max_objects}")
        centres = b["centres"][:n].cpu().numpy()
        if return_details:
            st = b["stats"].cpu().numpy()
            return centres, {"x": b["xy"][:n, 0].cpu().numpy(), "y": b["xy"][:n, 1].cpu().numpy(), "npix": b["npix"][:n].cpu().numpy(),
                             "globalback": st[0], "globalrms": st[1], "thresh": st[2]}
        return centres

    def plane(self, what):
        """intermediate plane of the last call as a torch tensor view copy: 'fg', 'conv', 'label', 'back', 'sigma', 'back_raw', 'sigma_raw'"""
        import torch

        from .. import _ffi

        b = self.last
        H, W = b["shape"]
        code = {"fg": 0, "conv": 1, "label": 2, "back": 3, "sigma": 4, "back_raw": 5, "sigma_raw": 6}[what]
        p = self._lib().dbv_detect_plane(C.c_void_p(b["base"]), H, W, self.max_objects, code)
        off = int(p) - b["scratch"].data_ptr()
        ny, nx = (H - 1) // 64 + 1, (W - 1) // 64 + 1
        shape = (H, W) if code < 3 else (ny, nx)
        nbytes = shape[0] * shape[1] * 4
        raw = b["scratch"][off : off + nbytes].clone()
        return raw.view(torch.int32 if code == 2 else torch.float32).reshape(shape)


def reduce_mesh_maps(maps, group=None):
    """tiled detection, exchange 1: every rank has written the statistics of ITS meshes into a (2, ny, nx) tensor preset to -inf;
    one all-reduce(MAX) completes the maps on every rank (a mesh computed by two ranks has the same bits on both)."""
    import torch.distributed as dist

    dist.all_reduce(maps, op=dist.ReduceOp.MAX, group=group)
    return maps


def merge_owned_objects(rows, flag, world, group=None):
    """tiled detection, exchange 2: ``rows`` = (k, 6) float64 tensor [order key, centre row, centre col, x, y, npix] of the objects this
    rank owns, ``flag`` != 0 when one of them reaches the rim of the rank's region.  Every rank learns every rank's (k, flag), then one
    padded all-gather; returns ``(merged (N, 6) ndarray in ascending order key — identical on every rank —, heads)`` or ``(None,
    heads)`` when any rank raised its flag (the caller then detects on the assembled field)."""
    import torch
    import torch.distributed as dist

    k = int(rows.shape[0])
    head = torch.tensor([k, int(flag)], dtype=torch.int64, device=rows.device)
    heads = [torch.empty_like(head) for _ in range(world)]
    dist.all_gather(heads, head, group=group)
    heads = torch.stack(heads).cpu().numpy()
    if heads[:, 1].any():
        return None, heads
    nmax = int(heads[:, 0].max())
    if nmax == 0:
        return np.zeros((0, 6)), heads
    mine = torch.zeros((nmax, 6), dtype=torch.float64, device=rows.device)
    if k:
        mine[:k] = rows
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine, group=group)
    allr = np.concatenate([parts[r][: int(heads[r, 0])].cpu().numpy() for r in range(world)])
    return allr[np.argsort(allr[:, 0], kind="stable")], heads


class TiledDeviceDetector(DeviceDetector):
    """The device detector on a field tiled over GPUs (one process per GPU; ``debvader_b200.parallel.LocalField``): every rank works
    on its owner tile + halo only and all ranks return the SAME global list of centres — the single-GPU list, bit for bit.

    Data path: mesh statistics of the rank's own meshes -> ONE all-reduce(MAX) of the (2, ny, nx) mesh maps (32 KB for a 4096^2
    field) -> background / filter / components on the region -> the objects whose last pixel lies in the owner tile -> one
    all-gather of (order key, row, col) per object, merged by the order key.  No field-sized transfer.  Falls back to detection
    on the assembled field (``parallel.gather_field``) when that is needed for exactness: tiles not aligned to the 64-px mesh
    grid, or an owned object reaching the rim of the region (footprints wider than the 30-px halo minus the filter's 3 px)."""

    accepts_local = True

    def __init__(self, group=None, **kw):
        super().__init__(**kw)
        self.group = group
        self.fallbacks = 0
        self._rbuf = {}

    @staticmethod
    def meshes_covered(field_size, world, halo):
        """every 64x64 mesh of the field lies wholly inside some rank's region"""
        from .. import parallel

        regs = parallel.region_bounds(field_size, world, halo)
        n = (field_size - 1) // 64 + 1
        for my in range(n):
            y0, y1 = 64 * my, min(64 * my + 64, field_size)
            for mx in range(n):
                x0, x1 = 64 * mx, min(64 * mx + 64, field_size)
                if not any(R0 <= y0 and y1 <= R1 and C0 <= x0 and x1 <= C1 for R0, R1, C0, C1 in regs):
                    return False
        return True

    def _region_buffers(self, F, RH, RW, dev):
        import torch

        from .. import _ffi

        key = (F, RH, RW, str(dev))
        b = self._rbuf.get(key)
        if b is None:
            M = self.max_objects
            nbytes = int(self._lib().dbv_detect_scratch_bytes_region(F, F, RH, RW, M))
            n = (F - 1) // 64 + 1
            b = {"scratch": torch.empty(nbytes + 256, dtype=torch.uint8, device=dev), "nbytes": nbytes,
                 "maps": torch.empty((2, n, n), dtype=torch.float32, device=dev),
                 "n": torch.zeros(1, dtype=torch.int32, device=dev), "flags": torch.zeros(4, dtype=torch.int32, device=dev),
                 "xy": torch.empty((M, 2), dtype=torch.float64, device=dev), "centres": torch.empty((M, 2), dtype=torch.float64, device=dev),
                 "npix": torch.empty(M, dtype=torch.int32, device=dev), "last": torch.empty(M, dtype=torch.int64, device=dev),
                 "stats": torch.zeros(4, dtype=torch.float32, device=dev)}
            b["base"] = b["scratch"].data_ptr() + (-b["scratch"].data_ptr()) % 256
            self._rbuf = {key: b}
        return b

    def __call__(self, field_image, local=None, return_details=False):
        import torch
        import torch.distributed as dist

        from .. import _ffi, parallel

        if local is None or local.world == 1:
            return super().__call__(local.data if (local is not None and field_image is None) else field_image, return_details=return_details)
        t = local.data if field_image is None else field_image
        if t.ndim == 4:
            t = t[0]
        self._require_device(t)
        if t.dtype not in (torch.float64, torch.float32) or not t.is_contiguous():
            raise ValueError("the tiled detector needs the rank's contiguous region tensor (float64 / float32)")
        F = local.field_size
        R0, R1, C0, C1 = local.region
        r0, r1, c0, c1 = local.tile
        RH, RW, Cn = (int(v) for v in t.shape)
        if (RH, RW) != (R1 - R0, C1 - C0):
            raise ValueError(f"region tensor {tuple(t.shape)} does not match the rank's region {local.region}")
        if not self.meshes_covered(F, local.world, local.halo):
            return self._on_assembled_field(local, field_image, return_details)
        b = self._region_buffers(F, RH, RW, t.device)
        lib = self._lib()
        with self._device_ctx(t):
            b["maps"].fill_(float("-inf"))
            self._check(lib.dbv_detect_meshes(_ffi.ptr(t), 1 if t.dtype == torch.float64 else 0, RH, RW, RW, Cn, self.band, R0, C0, F, F, self.max_objects,
                                             C.c_void_p(b["base"]), b["nbytes"], _ffi.ptr(b["maps"][0]), _ffi.ptr(b["maps"][1]), self._stream()))
            reduce_mesh_maps(b["maps"], self.group)
            self._check(lib.dbv_detect_objects(RH, RW, R0, C0, F, F, _ffi.ptr(b["maps"][0]), _ffi.ptr(b["maps"][1]), self.taps.ctypes.data_as(C.c_void_p),
                                              int(self.taps.shape[0]), int(self.taps.shape[1]), self.thresh, self.minarea, int(F / 2), int(F / 2),
                                              r0, r1, c0, c1, self.max_objects, C.c_void_p(b["base"]), b["nbytes"], _ffi.ptr(b["n"]), _ffi.ptr(b["xy"]),
                                              _ffi.ptr(b["centres"]), _ffi.ptr(b["npix"]), _ffi.ptr(b["last"]), _ffi.ptr(b["flags"]), _ffi.ptr(b["stats"]),
                                              self._stream()))
            nf = b["n"].cpu().numpy()[0], b["flags"].cpu().numpy()[0]  # one synchronisation
            k = min(int(nf[0]), self.max_objects)
            rows = torch.empty((k, 6), dtype=torch.float64, device=t.device)
            if k:
                rows[:, 0] = b["last"][:k].to(torch.float64)  # < 2^31: exact
                rows[:, 1:3] = b["centres"][:k]
                rows[:, 3:5] = b["xy"][:k]
                rows[:, 5] = b["npix"][:k].to(torch.float64)
            merged, _ = merge_owned_objects(rows, int(nf[1]) or int(nf[0] > self.max_objects), local.world, self.group)
        if merged is None:
            return self._on_assembled_field(local, field_image, return_details)
        centres = np.ascontiguousarray(merged[:, 1:3])
        if return_details:
            st = b["stats"].cpu().numpy()
            return centres, {"x": merged[:, 3], "y": merged[:, 4], "npix": merged[:, 5].astype(np.int32), "last": merged[:, 0].astype(np.int64),
                             "globalback": st[0], "globalrms": st[1], "thresh": st[2]}
        return centres

    def _on_assembled_field(self, local, field_image, return_details):
        """the exact but replicated path: every rank assembles the detection band of the whole field ON ITS DEVICE from the owner tiles
        (one all-gather of F x F values, 134 MB for a 4096^2 f64 field — no host transfer) and runs the single-GPU detector on it"""
        import torch
        import torch.distributed as dist

        from .. import parallel

        self.fallbacks += 1
        data = local.data if field_image is None else field_image
        tile = local.owner_tile(data)[..., self.band].contiguous()
        F = local.field_size
        tb = parallel.tile_bounds(F, local.world)
        mh, mw = max(b[1] - b[0] for b in tb), max(b[3] - b[2] for b in tb)
        pad = torch.zeros((mh, mw), dtype=tile.dtype, device=tile.device)
        pad[: tile.shape[0], : tile.shape[1]] = tile
        parts = [torch.empty_like(pad) for _ in range(local.world)]
        dist.all_gather(parts, pad, group=self.group)
        full = torch.empty((1, F, F, 1), dtype=tile.dtype, device=tile.device)
        for r, (r0, r1, c0, c1) in enumerate(tb):
            full[0, r0:r1, c0:c1, 0] = parts[r][: r1 - r0, : c1 - c0]
        del parts, pad
        return DeviceDetector.__call__(self, full, return_details=return_details, band=0)


_default_device_detector = None


def detect_objects_device(field_image, return_details=False):
    """backend="device" of detect_objects with a module-level DeviceDetector."""
    global _default_device_detector
    if _default_device_detector is None:
        _default_device_detector = DeviceDetector()
    return _default_device_detector(field_image, return_details=return_details)


def detect_objects(field_image, backend=None):
    """Detect objects on the r band (index 2); returns (row, col) offsets from the field centre (detection.py:5-56).

    backend: "sep" (the reference's CPU library), "device" (CUDA kernels), None = "device" for CUDA tensors, "sep" otherwise."""
    is_cuda = hasattr(field_image, "is_cuda") and field_image.is_cuda
    if backend is None:
        backend = "device" if is_cuda else "sep"
    if backend == "device":
        return detect_objects_device(field_image)
    if backend != "sep":
        raise ValueError(f"unknown detection backend {backend!r}")
    try:
        import sep
    except ImportError as e:  # pragma: no cover
        raise ImportError("detect_objects(backend='sep') needs the third-party `sep` package (SExtractor); use backend='device', "
                          "IterativeDeblendField(..., detector='device'), or pass galaxy_distances_to_center / a detector= callable") from e
    if hasattr(field_image, "detach"):
        field_image = field_image.detach().cpu().numpy()
    field_image = np.asarray(field_image).copy()
    field_size = field_image.shape[1]
    r_band = field_image[0, :, :, R_BAND].copy()
    bkg = sep.Background(r_band)
    objects = sep.extract(data=r_band - bkg, thresh=DETECT_THRESH, err=bkg.globalrms, deblend_cont=0.00001, deblend_nthresh=64,
                          minarea=MINAREA, filter_kernel=FILTER_KERNEL, filter_type="conv")
    out = [(np.round(-int(field_size / 2) + objects["y"][i]), np.round(-int(field_size / 2) + objects["x"][i]))
           for i in range(len(objects["y"]))]
    return np.array(out)
