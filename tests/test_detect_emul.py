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


def _tiled_worker(rank, world, port, so_path, q):
    """one gloo rank of the tiled detector: the emulated kernels for the device work, the PRODUCT's exchange functions
    (detect/detection.py: reduce_mesh_maps, merge_owned_objects) over torch.distributed for the two collectives"""
    import torch
    import torch.distributed as dist

    from debvader_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lib = C.CDLL(so_path)
        lib.dbv_detect_scratch_bytes_region.restype = C.c_int64
        lib.dbv_detect_scratch_bytes_region.argtypes = [C.c_int64] * 5
        lib.dbv_detect_meshes.restype = C.c_int
        lib.dbv_detect_meshes.argtypes = [C.c_void_p, C.c_int] + [C.c_int64] * 3 + [C.c_int, C.c_int] + [C.c_int64] * 5 + [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.dbv_detect_objects.restype = C.c_int
        lib.dbv_detect_objects.argtypes = [C.c_int64] * 6 + [C.c_void_p] * 3 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int] + [C.c_int64] * 5 + \
            [C.c_void_p, C.c_int64] + [C.c_void_p] * 8
        results = {}
        for tag, wide in (("plain", False), ("wide", True)):
            field = make_field(128, 14, seed=41)[0]
            if wide:
                yy, xx = np.mgrid[0:128, 0:128]
                field[0] += (40.0 * np.exp(-((xx - 66) ** 2 + (yy - 50) ** 2) / (2 * 12.0 ** 2)))[..., None]
            F, M = 128, 256
            assert det.TiledDeviceDetector.meshes_covered(F, world, 30)
            R0, R1, C0, C1 = parallel.region_bounds(F, world, 30)[rank]
            r0, r1, c0, c1 = parallel.tile_bounds(F, world)[rank]
            reg = np.ascontiguousarray(field[0, R0:R1, C0:C1])
            RH, RW = reg.shape[:2]
            nbytes = lib.dbv_detect_scratch_bytes_region(F, F, RH, RW, M)
            raw = np.zeros(nbytes + 256, np.uint8)
            base = raw.ctypes.data + (-raw.ctypes.data) % 256
            maps = torch.full((2, 2, 2), float("-inf"), dtype=torch.float32)
            assert lib.dbv_detect_meshes(reg.ctypes.data, 1, RH, RW, RW, 6, 2, R0, C0, F, F, M, base, nbytes, maps[0].data_ptr(), maps[1].data_ptr(), None) == 0
            det.reduce_mesh_maps(maps)
            assert torch.isfinite(maps).all()
            taps = det.normalised_taps()
            n, flags, stats = np.zeros(1, np.int32), np.zeros(4, np.int32), np.zeros(4, np.float32)
            xy, cen, npix, last = np.zeros((M, 2)), np.zeros((M, 2)), np.zeros(M, np.int32), np.zeros(M, np.int64)
            assert lib.dbv_detect_objects(RH, RW, R0, C0, F, F, maps[0].data_ptr(), maps[1].data_ptr(), taps.ctypes.data, 7, 7, 1.5, 4, 64, 64, r0, r1, c0, c1, M, base,
                                          nbytes, n.ctypes.data, xy.ctypes.data, cen.ctypes.data, npix.ctypes.data, last.ctypes.data, flags.ctypes.data,
                                          stats.ctypes.data, None) == 0
            k = int(n[0])
            rows = torch.from_numpy(np.concatenate([last[:k, None].astype(np.float64), cen[:k], xy[:k], npix[:k, None].astype(np.float64)], axis=1))
            merged, heads = det.merge_owned_objects(rows, int(flags[0]), world)
            results[tag] = (None if merged is None else merged.copy(), heads.copy())
        q.put((rank, results))
    finally:
        dist.destroy_process_group()


def test_tiled_detector_exchanges_over_gloo(emu):
    """world_size 2 over gloo: the N > 1 choreography of TiledDeviceDetector (all-reduce(MAX) of the mesh maps preset to -inf, counts +
    flags, padded all-gather, merge by the order key, the collective decision to take the assembled-field path) with the emulated
    kernels standing in for the GPU — every rank ends with the whole-field oracle's list"""
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
    c_ref, o = D.detect(make_field(128, 14, seed=41)[0], det.FILTER_KERNEL, return_details=True)
    assert len(c_ref) > 5
    for rank in (0, 1):
        merged, heads = got[rank]["plain"]
        assert heads[:, 1].sum() == 0 and heads[:, 0].sum() == len(c_ref)
        np.testing.assert_array_equal(merged[:, 0].astype(np.int64), o["last"])
        np.testing.assert_array_equal(merged[:, 1:3], c_ref)
        np.testing.assert_array_equal(merged[:, 3], o["x"])
        np.testing.assert_array_equal(merged[:, 5].astype(np.int64), o["npix"])
        merged_w, heads_w = got[rank]["wide"]
        assert merged_w is None and heads_w[:, 1].any()  # both ranks decide together to detect on the assembled field


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
