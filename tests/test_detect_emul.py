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
