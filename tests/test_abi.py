"""The C-ABI library loads without a GPU and exports every symbol include/debvader_b200.h declares."""
import ctypes
import os
import re

import pytest
import torch

from debvader_b200 import _build, _ffi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _build.build()
    return _ffi.lib()


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "debvader_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dbv_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported(lib):
    names = declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/debvader_b200.h but not exported"
    assert set(names) == set(_ffi.EXPORTS), "ctypes signatures out of sync with the header"


def test_product_library_has_no_environment_switches_or_debug_exports(lib):
    """kernel A/B switches, timing ablations and the descriptor probes live in the ablation build
    (libdebvader_b200_ablate.so, -DDBV_ABLATE) only."""
    data = open(_ffi.LIB_PATH, "rb").read()
    names = set(re.findall(rb"DBV_[A-Z0-9_]{3,}", data))
    assert names <= {b"DBV_VERBOSE"}, names  # DBV_VERBOSE only prints the plans the autotuner picked
    assert not hasattr(lib, "dbv_probe") and not hasattr(lib, "dbv_halo_counters")
    _build.build(ablate=True)
    dbg = ctypes.CDLL(_ffi.ABLATE_LIB_PATH)
    assert hasattr(dbg, "dbv_probe") and hasattr(dbg, "dbv_halo_counters") and dbg.dbv_abi_version() == _ffi.ABI_VERSION
    assert b"DBV_HALO_SKIP" in open(_ffi.ABLATE_LIB_PATH, "rb").read()


def test_abi_version(lib):
    assert lib.dbv_abi_version() == _ffi.ABI_VERSION == 6
    assert lib.dbv_mse_scratch_bytes() > 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_gpu(lib):
    ctx = ctypes.c_void_p()
    rc = lib.dbv_create(ctypes.byref(ctx), 0, 0, 0)
    assert rc < 0
    assert b"no CUDA device" in lib.dbv_last_error() or b"CPU" in lib.dbv_last_error()
    from debvader_b200.model.model import load_deblender

    with pytest.raises(RuntimeError):
        load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random")


def test_non_dc2_architecture_is_refused():
    from debvader_b200.model.model import load_deblender

    with pytest.raises(NotImplementedError):
        load_deblender("dc2", (64, 64, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random")


def test_host_pipeline_schedule(lib, monkeypatch):
    """piece sizes of dbv_deblend_host (pure host function): they cover the batch exactly, never exceed the context's
    chunk, start small, grow no faster than the copies keep up with the compute and end with one short piece."""
    monkeypatch.delenv("DBV_HOST_PIECE", raising=False)

    def sched(B, chunk=4096):
        buf = (ctypes.c_int64 * 4096)()
        n = lib.dbv_host_schedule(B, chunk, ctypes.cast(buf, ctypes.c_void_p), 4096)
        assert 0 <= n <= 4096
        return list(buf[:n])

    assert sched(0) == []
    assert sched(1) == [1] and sched(300) == [300]
    assert sched(512) == [256, 256]
    assert sched(4096) == [256, 544, 896, 1344, 800, 256]
    for chunk in (64, 256, 1024, 4096):
        for B in list(range(1, 3000, 7)) + [4096, 8192, 100000]:
            s = sched(B, chunk)
            assert sum(s) == B and min(s) > 0 and max(s) <= chunk, (B, chunk, s)
            for a, b in zip(s, s[1:-1]):  # growth bound of the ramp (the last piece is the short one)
                assert b <= 5 * a // 4 + 224 or b <= a, (B, chunk, s)
    big = sched(100000)
    assert big[0] == 256 and big[-1] <= 320 and max(big) == 1376
    monkeypatch.setenv("DBV_HOST_PIECE", "1024")  # a tuning knob of the ablation build only: the product library ignores it
    assert sched(4096) == [256, 544, 896, 1344, 800, 256]
    assert lib.dbv_host_schedule(-1, 4096, None, 0) < 0
