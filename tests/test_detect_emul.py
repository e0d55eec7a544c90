"""CPU: the kernels of csrc/detect_kernels.cu compiled as host C++ (-DDBV_EMULATE: CUDA threads as std::threads, __syncthreads as a
barrier; tools/detect_emul/detect_emul.h) and run through the SAME dbv_detect entry point, compared bit for bit with
oracle/detect_numpy.py.  This checks the kernels' indexing, barriers and arithmetic order in the container, where there is no GPU;
it says nothing about the GPU build itself — tests/test_gpu_detect.py runs the library on the B200."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from debvader_b200.detect import detection as det
from oracle import detect_numpy as D
from tests.test_detect_oracle import make_field

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tools", "detect_emul")


@pytest.fixture(scope="module")
def emu():
    if not shutil.which("g++"):
        pytest.skip("g++ not available")
    out = os.path.join(EMU, "_build", "libdetect_emul.so")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    src = os.path.join(ROOT, "debvader_b200", "csrc", "detect_kernels.cu")
    subprocess.check_call(["g++", "-O1", "-std=c++20", "-ffp-contract=off", "-DDBV_EMULATE", "-x", "c++", src, "-I", EMU, "-shared", "-fPIC",
                           "-o", out, "-lpthread"])
    lib = C.CDLL(out)
    lib.dbv_detect_scratch_bytes.restype = C.c_int64
    lib.dbv_detect_scratch_bytes.argtypes = [C.c_int64] * 3
    lib.dbv_detect.restype = C.c_int
    lib.dbv_detect.argtypes = [C.c_void_p, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int,
                               C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_int64] + [C.c_void_p] * 6
    lib.dbv_detect_plane.restype = C.c_void_p
    lib.dbv_detect_plane.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int]
    return lib


def run_emulated(lib, field, max_objects=512):
    f = np.ascontiguousarray(field[0])
    H, W, Cn = f.shape
    nbytes = lib.dbv_detect_scratch_bytes(H, W, max_objects)
    raw = np.zeros(nbytes + 256, np.uint8)
    base = raw.ctypes.data + (-raw.ctypes.data) % 256
    n = np.zeros(1, np.int32)
    xy = np.zeros((max_objects, 2))
    cen = np.zeros((max_objects, 2))
    npix = np.zeros(max_objects, np.int32)
    stats = np.zeros(4, np.float32)
    taps = det.normalised_taps()
    rc = lib.dbv_detect(f.ctypes.data, 1 if f.dtype == np.float64 else 0, H, W, W, Cn, 2, taps.ctypes.data, 7, 7, 1.5, 4, int(H / 2), int(W / 2), max_objects,
                        base, nbytes, n.ctypes.data, xy.ctypes.data, cen.ctypes.data, npix.ctypes.data, stats.ctypes.data, None)
    assert rc == 0
    ny, nx = (H - 1) // 64 + 1, (W - 1) // 64 + 1

    def plane(code):
        p = lib.dbv_detect_plane(base, H, W, max_objects, code)
        shape = (H, W) if code < 3 else (ny, nx)
        a = np.frombuffer((C.c_char * (shape[0] * shape[1] * 4)).from_address(p), np.int32 if code == 2 else np.float32).reshape(shape)
        return a.copy()

    k = int(n[0])
    return {"n": k, "x": xy[:k, 0], "y": xy[:k, 1], "centres": cen[:k], "npix": npix[:k], "stats": stats, "fg": plane(0), "conv": plane(1),
            "label": plane(2), "back": plane(3), "sigma": plane(4), "back_raw": plane(5), "sigma_raw": plane(6), "_keep": raw}


@pytest.mark.parametrize("case", ["dc2_259", "rect_150x200_f32", "one_mesh_60x50"])
def test_emulated_kernels_match_the_oracle_bit_for_bit(emu, case, golden_dir):
    if case == "dc2_259":
        field = np.load(os.path.join(golden_dir, "dc2_field2.npz"))["field"]
    elif case == "rect_150x200_f32":
        field = make_field(200, 25, seed=21, gradient=0.03)[0][:, :150].astype(np.float32)
    else:
        field = make_field(64, 2, seed=22)[0][:, :60, :50]
    c_ref, o = D.detect(field, det.FILTER_KERNEL, return_details=True)
    e = run_emulated(emu, field)
    rep = {k: int((e[k] != o[k]).sum()) for k in ("back_raw", "sigma_raw", "back", "sigma", "fg", "conv")}
    rep["stats"] = int(e["stats"][0] != o["globalback"]) + int(e["stats"][1] != o["globalrms"]) + int(e["stats"][2] != o["thresh"])
    rep["mask"] = int(((e["label"] >= 0) != (o["conv"] > o["thresh"])).sum())
    rep["n"] = int(e["n"] != len(c_ref))
    if not rep["n"]:
        rep.update(npix=int((e["npix"] != o["npix"]).sum()), x=int((e["x"] != o["x"]).sum()), y=int((e["y"] != o["y"]).sum()),
                   centres=int((e["centres"] != c_ref).sum()))
    assert not any(rep.values()), rep
    assert e["n"] > 0


def run_emulated_tiled(lib, field, world, max_objects=512, halo=30):
    """the field split into owner tiles + halo as debvader_b200.parallel does, one emulated 'rank' after the other:
    dbv_detect_meshes per rank -> max-reduce of the mesh maps -> dbv_detect_objects per rank -> objects merged by their order key"""
    from debvader_b200 import parallel

    lib.dbv_detect_scratch_bytes_region.restype = C.c_int64
    lib.dbv_detect_scratch_bytes_region.argtypes = [C.c_int64] * 5
    lib.dbv_detect_meshes.restype = C.c_int
    lib.dbv_detect_meshes.argtypes = [C.c_void_p, C.c_int] + [C.c_int64] * 3 + [C.c_int, C.c_int] + [C.c_int64] * 5 + [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.dbv_detect_objects.restype = C.c_int
    lib.dbv_detect_objects.argtypes = [C.c_int64] * 6 + [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_int64] * 5 + \
        [C.c_void_p, C.c_int64] + [C.c_void_p] * 8
    F = field.shape[1]
    Cn = field.shape[3]
    ny = nx = (F - 1) // 64 + 1
    taps = det.normalised_taps()
    regions, tiles = parallel.region_bounds(F, world, halo), parallel.tile_bounds(F, world)
    ranks = []
    maps = np.full((2, ny, nx), -np.inf, np.float32)
    for (R0, R1, C0, C1) in regions:
        reg = np.ascontiguousarray(field[0, R0:R1, C0:C1])
        RH, RW = reg.shape[:2]
        nbytes = lib.dbv_detect_scratch_bytes_region(F, F, RH, RW, max_objects)
        raw = np.zeros(nbytes + 256, np.uint8)
        base = raw.ctypes.data + (-raw.ctypes.data) % 256
        mine = np.full((2, ny, nx), -np.inf, np.float32)
        rc = lib.dbv_detect_meshes(reg.ctypes.data, 1 if reg.dtype == np.float64 else 0, RH, RW, RW, Cn, 2, R0, C0, F, F, max_objects, base, nbytes,
                                   mine[0].ctypes.data, mine[1].ctypes.data, None)
        assert rc == 0
        maps = np.maximum(maps, mine)  # the all-reduce(MAX) of the tiled detector
        ranks.append((reg, raw, base, nbytes, RH, RW, R0, C0))
    assert np.isfinite(maps).all(), "a mesh lies in no rank's region"
    objs, flagged = [], 0
    for (reg, raw, base, nbytes, RH, RW, R0, C0), (r0, r1, c0, c1) in zip(ranks, tiles):
        n = np.zeros(1, np.int32)
        xy = np.zeros((max_objects, 2))
        cen = np.zeros((max_objects, 2))
        npix = np.zeros(max_objects, np.int32)
        last = np.zeros(max_objects, np.int64)
        flags = np.zeros(4, np.int32)
        stats = np.zeros(4, np.float32)
        rc = lib.dbv_detect_objects(RH, RW, R0, C0, F, F, maps[0].ctypes.data, maps[1].ctypes.data, taps.ctypes.data, 7, 7, 1.5, 4, int(F / 2), int(F / 2),
                                    r0, r1, c0, c1, max_objects, base, nbytes, n.ctypes.data, xy.ctypes.data, cen.ctypes.data, npix.ctypes.data,
                                    last.ctypes.data, flags.ctypes.data, stats.ctypes.data, None)
        assert rc == 0
        k = int(n[0])
        flagged += int(flags[0])
        objs += [(int(last[i]), xy[i, 0], xy[i, 1], int(npix[i]), cen[i, 0], cen[i, 1]) for i in range(k)]
    objs.sort(key=lambda o: o[0])
    return objs, flagged, stats


@pytest.mark.parametrize("world", [2, 8])
def test_emulated_tiled_detection_equals_the_whole_field(emu, world):
    """owner tile + 30-px halo per rank, mesh maps max-reduced, objects owned by the tile of their last pixel: the merged list is the
    whole-field list, bit for bit (256^2 field: tiles of 128 / 64 px, i.e. mesh-aligned)"""
    field = make_field(256, 45, seed=31, gradient=0.02)[0]
    c_ref, o = D.detect(field, det.FILTER_KERNEL, return_details=True)
    objs, flagged, stats = run_emulated_tiled(emu, field, world)
    assert flagged == 0
    assert stats[1] == o["globalrms"] and stats[2] == o["thresh"]
    assert [t[0] for t in objs] == list(o["last"])
    assert [t[3] for t in objs] == list(o["npix"])
    np.testing.assert_array_equal(np.array([t[1] for t in objs]), o["x"])
    np.testing.assert_array_equal(np.array([t[2] for t in objs]), o["y"])
    np.testing.assert_array_equal(np.array([[t[4], t[5]] for t in objs]), c_ref)


def test_emulated_tiled_detection_flags_an_object_that_leaves_the_region(emu):
    field = make_field(256, 10, seed=32)[0]
    yy, xx = np.mgrid[0:256, 0:256]
    field[0] += (40.0 * np.exp(-((xx - 131) ** 2 + (yy - 100) ** 2) / (2 * 14.0 ** 2)))[..., None]  # footprint ~ 60 px across the tile edge at x = 128
    objs, flagged, _ = run_emulated_tiled(emu, field, 2)
    assert flagged >= 1


def _declare(lib):
    """ctypes signatures of the emulated library = include/debvader_b200.h"""
    i64, vp = C.c_int64, C.c_void_p
    lib.dbv_detect_scratch_bytes.restype = i64
    lib.dbv_detect_scratch_bytes.argtypes = [i64] * 3
    lib.dbv_detect_scratch_bytes_region.restype = i64
    lib.dbv_detect_scratch_bytes_region.argtypes = [i64] * 5
    lib.dbv_detect.restype = C.c_int
    lib.dbv_detect.argtypes = [vp, C.c_int, i64, i64, i64, C.c_int, C.c_int, vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int, i64, vp, i64] + [vp] * 6
    lib.dbv_detect_meshes.restype = C.c_int
    lib.dbv_detect_meshes.argtypes = [vp, C.c_int] + [i64] * 3 + [C.c_int, C.c_int] + [i64] * 5 + [vp, i64, vp, vp, vp]
    lib.dbv_detect_objects.restype = C.c_int
    lib.dbv_detect_objects.argtypes = [i64] * 6 + [vp] * 3 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int] + [i64] * 5 + [vp, i64] + [vp] * 8
    lib.dbv_detect_plane.restype = vp
    lib.dbv_detect_plane.argtypes = [vp, i64, i64, i64, C.c_int]
    return lib


def _emulated_detector_class(lib):
    """the PRODUCT's TiledDeviceDetector with its five device hooks pointed at the emulated library and CPU tensors: everything else —
    buffers, the two phases, synchronisation, both exchanges, the collective fall-back decision, the assembled-band path — is product code"""
    import contextlib

    class EmuTiled(det.TiledDeviceDetector):
        def _lib(self):
            return lib

        def _check(self, rc):
            assert rc == 0, rc
            return rc

        def _device_ctx(self, t):
            return contextlib.nullcontext()

        def _stream(self):
            return None

        def _require_device(self, t):
            pass

    return EmuTiled


def _tiled_worker(rank, world, port, so_path, q):
    """one gloo rank: TiledDeviceDetector itself (detect/detection.py) on this rank's LocalField, CPU tensors, emulated kernels"""
    import torch
    import torch.distributed as dist

    from debvader_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Emu = _emulated_detector_class(_declare(C.CDLL(so_path)))
        results = {}
        for tag, wide in (("plain", False), ("wide", True)):
            field = make_field(128, 14, seed=41)[0]
            if wide:
                yy, xx = np.mgrid[0:128, 0:128]
                field[0] += (40.0 * np.exp(-((xx - 66) ** 2 + (yy - 50) ** 2) / (2 * 12.0 ** 2)))[..., None]
            R0, R1, C0, C1 = parallel.region_bounds(128, world, 30)[rank]
            local = parallel.LocalField(torch.from_numpy(np.ascontiguousarray(field[:, R0:R1, C0:C1])), 128, rank, world)
            d = Emu(device="cpu", max_objects=256)
            centres, info = d(local.data, local, return_details=True)
            again = d(local.data, local)  # buffers are reused
            assert np.array_equal(again, centres)
            results[tag] = (centres, info["x"], info["npix"], d.fallbacks)
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_tiled_detector_class_over_gloo(emu):
    """world_size 2 over gloo: the product's TiledDeviceDetector end to end — all-reduce(MAX) of the mesh maps preset to -inf, counts +
    flags, padded all-gather, merge by the order key, and, for a footprint wider than the halo, the collective decision to assemble the
    detection band on every rank and detect on it — with the emulated kernels standing in for the GPU.  Every rank must end with the
    whole-field oracle's list."""
    import socket

    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    so = os.path.join(EMU, "_build", "libdetect_emul.so")
    procs = [ctx.Process(target=_tiled_worker, args=(r, 2, port, so, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=600) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    field = make_field(128, 14, seed=41)[0]
    c_ref, o = D.detect(field, det.FILTER_KERNEL, return_details=True)
    yy, xx = np.mgrid[0:128, 0:128]
    wide = field.copy()
    wide[0] += (40.0 * np.exp(-((xx - 66) ** 2 + (yy - 50) ** 2) / (2 * 12.0 ** 2)))[..., None]
    w_ref, w_o = D.detect(wide, det.FILTER_KERNEL, return_details=True)
    assert len(c_ref) > 5
    for rank in (0, 1):
        centres, x, npix, fallbacks = got[rank]["plain"]
        assert fallbacks == 0
        np.testing.assert_array_equal(centres, c_ref)
        np.testing.assert_array_equal(x, o["x"])
        np.testing.assert_array_equal(npix, o["npix"])
        centres, x, npix, fallbacks = got[rank]["wide"]
        assert fallbacks == 2  # both calls took the assembled-band path, on both ranks
        np.testing.assert_array_equal(centres, w_ref)
        np.testing.assert_array_equal(x, w_o["x"])


def test_single_gpu_detector_class_on_the_emulated_library(emu):
    """DeviceDetector's own host logic (buffers, run / call, plane views, host arrays) with the emulated kernels"""
    import torch

    Emu = _emulated_detector_class(_declare(emu))
    field = make_field(100, 8, seed=43)[0]
    c_ref, o = D.detect(field, det.FILTER_KERNEL, return_details=True)
    d = Emu(device="cpu", max_objects=64)
    c, info = d(torch.from_numpy(field), return_details=True)
    np.testing.assert_array_equal(c, c_ref)
    np.testing.assert_array_equal(info["y"], o["y"])
    assert np.array_equal(d.plane("conv").numpy(), o["conv"]) and np.array_equal(d.plane("back").numpy(), o["back"])
    np.testing.assert_array_equal(d(field), c_ref)  # a host array is "uploaded" to the detector's device
    with pytest.raises(ValueError):
        d(torch.from_numpy(field[:, :, :, :2]))  # no band 2
    small = Emu(device="cpu", max_objects=2)
    with pytest.raises(RuntimeError, match="more than max_objects"):
        small(torch.from_numpy(field))


@pytest.mark.parametrize("case", ["zeros", "constant", "tiny_12x9", "lone_pixels", "step_background"])
def test_emulated_kernels_on_degenerate_fields(emu, case):
    """edge cases: zero / constant fields (sigma = 0: one histogram level, threshold 0, nothing above it), a field smaller than the filter
    footprint of its own corners, detections below minarea, a background step across meshes — no hang, no NaN, same answer as the oracle"""
    rng = np.random.default_rng(51)
    if case == "zeros":
        field = np.zeros((1, 70, 70, 6))
    elif case == "constant":
        field = np.full((1, 70, 70, 6), 3.25)
    elif case == "tiny_12x9":
        field = rng.normal(0, 0.03, (1, 12, 9, 6))
        field[0, 4:8, 3:6] += 1.0
    elif case == "lone_pixels":
        field = rng.normal(0, 0.03, (1, 70, 70, 6))
        field[0, 10, 10] += 0.5   # the filter spreads it, but too faintly for 4 pixels above the threshold
        field[0, 40:43, 40:43] += 2.0
    else:
        field = rng.normal(0, 0.03, (1, 130, 130, 6))
        field[0, :, 64:] += 0.5
    c_ref, o = D.detect(field, det.FILTER_KERNEL, return_details=True)
    e = run_emulated(emu, field)
    assert e["n"] == len(c_ref)
    assert np.array_equal(e["conv"], o["conv"]) and np.array_equal(e["fg"], o["fg"]) and np.array_equal(e["back"], o["back"])
    assert e["stats"][1] == o["globalrms"] and np.isfinite(e["stats"]).all()
    np.testing.assert_array_equal(e["centres"], c_ref)
    if case in ("zeros", "constant"):
        assert e["n"] == 0
    if case == "lone_pixels":
        assert e["n"] >= 1
