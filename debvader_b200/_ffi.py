"""ctypes binding of libdebvader_b200.so (the C-ABI in include/debvader_b200.h).

There is no CPU path: if the library is missing or a call fails, this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DEBVADER_B200_LIB") or os.path.join(_HERE, "libdebvader_b200.so")  # env override: A/B builds of the same source

ABI_VERSION = 6  # must equal DBV_ABI_VERSION in include/debvader_b200.h
PREC = {"fp32": 0, "bf16": 1, "bf16x3": 2, "fp16x3": 3, "mixed": 4, "fp32tc": 5}
F32, F64 = 0, 1

_lib = None
_lock = threading.Lock()

c_i64 = C.c_int64
c_vp = C.c_void_p


class DbvError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"debvader_b200 error {code}: {msg}")
        self.code = code


def _declare(lib):
    sig = {
        "dbv_abi_version": (C.c_int, []),
        "dbv_last_error": (C.c_char_p, []),
        "dbv_create": (C.c_int, [C.POINTER(c_vp), C.c_int, C.c_int, c_i64]),
        "dbv_destroy": (C.c_int, [c_vp]),
        "dbv_set_weights": (C.c_int, [c_vp, C.c_char_p, c_vp, C.POINTER(c_i64), C.c_int]),
        "dbv_finalize_weights": (C.c_int, [c_vp]),
        "dbv_encode": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp]),
        "dbv_latent": (C.c_int, [c_vp, c_vp, c_vp, C.c_uint64, C.c_int, c_i64, c_i64, c_vp, c_vp, c_vp, c_vp]),
        "dbv_decode": (C.c_int, [c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
        "dbv_deblend": (C.c_int, [c_vp, c_vp, c_i64, c_vp, C.c_uint64, C.c_int, c_vp, c_vp, c_vp, c_vp]),
        "dbv_deblend_host": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_vp, C.c_uint64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "dbv_host_schedule": (c_i64, [c_i64, c_i64, c_vp, c_i64]),
        "dbv_extract": (C.c_int, [c_vp, C.c_int, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_i64, C.c_int, c_vp, C.c_int, c_vp]),
        "dbv_window_axpy": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, C.c_int, c_vp, c_vp, c_vp, c_i64, C.c_int, C.c_double, c_vp]),
        "dbv_window_axpy_ex": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, C.c_int, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_i64, C.c_int, C.c_double, c_vp]),
        "dbv_window_axpy_scratch_bytes": (c_i64, [c_i64, c_i64]),
        "dbv_window_axpy_rect": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_i64, C.c_int, c_vp, C.c_int, C.c_int, c_vp, c_vp, c_i64, C.c_int, C.c_double,
                                           c_vp, c_i64, c_vp]),
        "dbv_sqdiff_sum_rect": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_i64, c_i64, c_i64, c_vp, c_vp, c_i64, c_vp]),
        "dbv_spline_extent": (C.c_int, [C.c_int, C.c_int]),
        "dbv_spline_scratch_doubles": (c_i64, [c_i64, C.c_int, C.c_int, C.c_int]),
        "dbv_spline_place": (C.c_int, [c_vp, C.c_int, c_i64, C.c_int, C.c_int, c_i64, C.c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, C.c_int, c_vp, c_vp, c_vp]),
        "dbv_band_sumsq": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_vp, c_vp, c_i64, c_vp]),
        "dbv_shift_objective": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, C.c_double, c_vp, c_vp]),
        "dbv_shift_objective_batch": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_vp, C.c_int, c_vp, c_vp, c_i64, C.c_double, c_vp, c_vp]),
        "dbv_position_objective": (C.c_int, [c_vp, c_i64, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                             C.c_double, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "dbv_center_mse": (C.c_int, [c_vp, C.c_int, c_vp, c_i64, C.c_int, C.c_int, C.c_int, C.c_int, c_vp, c_vp]),
        "dbv_mse": (C.c_int, [c_vp, c_vp, C.c_int, c_i64, c_vp, c_vp, c_i64, c_vp]),
        "dbv_mse_scratch_bytes": (c_i64, []),
        "dbv_fp16_overflow": (C.c_int, [c_vp, C.c_int]),
        "dbv_launch_count": (c_i64, [c_vp]),
        "dbv_global_launch_count": (c_i64, []),
        "dbv_set_profiling": (C.c_int, [c_vp, C.c_int]),
        "dbv_layer_times": (C.c_int, [c_vp, C.c_int, c_vp, c_vp]),
        "dbv_layer_kernel": (C.c_int, [c_vp, C.c_char_p, c_vp, C.c_int]),
        "dbv_detect_scratch_bytes": (c_i64, [c_i64, c_i64, c_i64]),
        "dbv_detect": (C.c_int, [c_vp, C.c_int, c_i64, c_i64, c_i64, C.c_int, C.c_int, c_vp, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                 c_i64, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp]),
        "dbv_detect_plane": (c_vp, [c_vp, c_i64, c_i64, c_i64, C.c_int]),
        "dbv_detect_plane_region": (c_vp, [c_vp, c_i64, c_i64, c_i64, c_i64, c_i64, C.c_int]),
        "dbv_detect_scratch_bytes_region": (c_i64, [c_i64] * 5),
        "dbv_detect_meshes": (C.c_int, [c_vp, C.c_int, c_i64, c_i64, c_i64, C.c_int, C.c_int, c_i64, c_i64, c_i64, c_i64, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
        "dbv_detect_objects": (C.c_int, [c_i64] * 6 + [c_vp] * 3 + [C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int] + [c_i64] * 5 + [c_vp, c_i64] + [c_vp] * 8),
        "dbv_debug_activation": (C.c_int, [c_vp, C.c_char_p, c_i64, c_vp, c_vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    return sig


EXPORTS = None


def lib():
    """Load the shared library (once).  Raises if it has not been built."""
    global _lib, EXPORTS
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise ImportError(
                        f"{LIB_PATH} is missing. Build it with `python -m debvader_b200._build` "
                        "(needs nvcc); debvader_b200 has no CPU fallback."
                    )
                l = C.CDLL(LIB_PATH)
                EXPORTS = _declare(l)
                if l.dbv_abi_version() != ABI_VERSION:
                    raise ImportError("libdebvader_b200.so ABI version mismatch; rebuild it")
                _lib = l
    return _lib


_dbg = None
ABLATE_LIB_PATH = os.path.join(_HERE, "libdebvader_b200_ablate.so")


def debug_lib():
    """The ablation build (-DDBV_ABLATE): DBV_* environment switches, dbv_probe, dbv_halo_counters
    (include/debvader_b200_debug.h).  A separate shared object with its own state; never used by the product path."""
    global _dbg
    if _dbg is None:
        with _lock:
            if _dbg is None:
                if not os.path.exists(ABLATE_LIB_PATH):
                    raise ImportError(f"{ABLATE_LIB_PATH} is missing. Build it with `python -m debvader_b200._build --ablate`.")
                l = C.CDLL(ABLATE_LIB_PATH)
                _declare(l)
                l.dbv_probe.restype = C.c_int
                l.dbv_probe.argtypes = [C.c_int, c_vp, c_vp, c_vp, C.c_int, C.c_int, C.c_int, c_vp]
                l.dbv_halo_counters.restype = C.c_int
                l.dbv_halo_counters.argtypes = [c_vp, C.c_int]
                _dbg = l
    return _dbg


def check(rc: int, l=None) -> int:
    if rc < 0:
        raise DbvError(rc, (l or lib()).dbv_last_error().decode("utf-8", "replace"))
    return rc


def ptr(t):
    """device/host pointer of a torch tensor (None -> NULL)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
