#!/usr/bin/env python
"""clock64 breakdown of the resident-halo kernel per layer (ablation build of the library).

    DEBVADER_B200_LIB=debvader_b200/libdebvader_b200_ablate.so python tools/halo_clocks.py [precision] [stamps]

For every halo layer: cycles per CTA spent by the MMA-issuing thread (waiting for a halo band / for a free accumulator
slot / issuing), by the TMA thread and by the epilogue warps (waiting for accumulators, tcgen05.ld, alpha loads, math +
stores), plus the MMA count, so that cycles per MMA and per item can be read directly.  Counters are summed over CTAs
(148) and epilogue warps (8) by the kernel; this script divides them back.
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("DEBVADER_B200_LIB", os.path.join(ROOT, "debvader_b200", "libdebvader_b200_ablate.so"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

from debvader_b200 import _ffi  # noqa: E402
from debvader_b200.model.model import load_deblender  # noqa: E402

NAMES = ["enc_conv1", "enc_conv2", "enc_conv3", "enc_conv4", "enc_conv5", "enc_conv6", "enc_conv7", "enc_conv8", "enc_dense", "dec_dense1",
         "dec_dense2", "dec_convT1", "dec_convT2", "dec_convT3", "dec_convT4", "dec_convT5", "dec_convT6", "dec_convT7", "dec_convT8", "dec_head"]


def main():
    precision = sys.argv[1] if len(sys.argv) > 1 else "mixed"
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    lib = _ffi.lib()
    assert hasattr(lib, "dbv_halo_counters"), "needs the ablation build (DEBVADER_B200_LIB=.../libdebvader_b200_ablate.so)"
    lib.dbv_halo_counters.restype = ctypes.c_int
    lib.dbv_halo_counters.argtypes = [ctypes.c_void_p, ctypes.c_int]
    net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights="random:1234", precision=precision)
    x = torch.randn((B, 59, 59, 6), device="cuda") * 0.3
    mean = torch.empty_like(x)
    std = torch.empty_like(x)
    for _ in range(3):
        net.deblend_into(x, mean, std)
    torch.cuda.synchronize()
    lib.dbv_halo_counters(None, 1)
    net.set_profiling(True)
    reps = 5
    for _ in range(reps):
        net.deblend_into(x, mean, std)
    torch.cuda.synchronize()
    times = dict(net.layer_times())
    buf = (ctypes.c_ulonglong * (24 * 16))()
    _ffi.check(lib.dbv_halo_counters(ctypes.cast(buf, ctypes.c_void_p), 0))
    c = np.array(list(buf), dtype=np.float64).reshape(24, 16)
    out = {}
    for li, name in enumerate(NAMES):
        r = c[li]
        if r[14] == 0:
            continue
        ctas = r[14] / reps  # CTAs per launch
        per = lambda v: v / r[14]  # cycles per CTA per launch
        out[name] = {
            "ms": round(times.get(name, 0.0), 4), "ctas": ctas,
            "mma_total": per(r[0]), "mma_wait_band": per(r[1]), "mma_wait_slot": per(r[2]), "mma_issue": per(r[3]),
            "mmas_per_cta": per(r[4]), "units_per_cta": per(r[5]), "cycles_per_mma_issue_only": r[3] / max(r[4], 1),
            "cycles_per_mma_total": r[0] / max(r[4], 1),
            "tma_total": per(r[6]), "tma_wait_free_buffer": per(r[7]),
            "epi_total_per_warp": per(r[8]) / 8, "epi_wait_acc": per(r[9]) / 8, "epi_tmem_ld": per(r[10]) / 8, "epi_alpha": per(r[11]) / 8,
            "epi_math_store": per(r[12]) / 8, "epi_release": per(r[15]) / 8, "items_per_warp": per(r[13]) / 8,
            "epi_cycles_per_item": (r[10] + r[11] + r[12]) / max(r[13], 1),
        }
    print(json.dumps({"precision": precision, "stamps": B, "layers": out}, indent=1))
    net.close()


if __name__ == "__main__":
    main()
