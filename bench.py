#!/usr/bin/env python
"""bench.py — headline benchmark of the debvader hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision mixed|bf16x3|fp16x3|bf16|fp32]

A "step" is one pass of the hot path (encode -> latent -> decode of BASELINE cfg 2: a batch of
4096 synthetic 59x59x6 stamps, 342 MB of fp32 input per step, i.e. larger than the 126 MB L2, so
consecutive steps never find their input in cache) on each GPU.  One process per GPU (torchrun for
N>1), stamps shard with no data-path collective -> weak scaling.  Prints ONE JSON line (rank 0):

  value      deblended stamps/s, whole job, inputs resident in HBM, timed with CUDA events on the
             launching stream between barrier+synchronize, max over ranks
  e2e        the same metric through the public call deblend(net, images) with HOST buffers:
             H2D of the step's input from pinned memory + D2H of the mean (what deblend() returns as
             an ndarray; the stddev stays on the device behind the returned distribution object)
             inside the timed region (wall clock around the synchronous call)
  roofline   the dominant kernel (the __global__ function with the largest share of the step, as the ncu launch
             list groups them) against the measured tensor-core peak; the slowest single layer is reported next to it
  cpu_baseline  the CPU oracle (torch-CPU restatement of the reference model — a stand-in, NOT
             TensorFlow, which is not installable here) on a bounded sample of the same workload
  field      extras: extraction / scatter kernels in GB/s against measured HBM bandwidth, and ms per
             4096^2 field with 2000 sources

--impl reference times the reference's CPU path stand-in (oracle port) on the host cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "deblended stamps/sec (encode+decode)"
UNIT = "stamps/s"
BATCH = 4096
CFG = ("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3])
STAMP_ELTS = 59 * 59 * 6
# which hand-written kernel runs each layer in the tensor-core precisions (csrc/api.cu: kTc, consumes/has_pair/has_halo)
KERNEL_OF = {**{k: "tc_halo_kernel" for k in ("enc_conv1", "enc_conv2", "enc_conv3", "dec_convT7")},
             **{k: "tc_halo2_kernel (cta_group::2 resident halo) or tc_halo_kernel, whichever the plan tuner timed faster" for k in ("dec_convT6", "dec_convT8", "dec_head")},
             **{k: "tc_pairh_kernel (cta_group::2, halo box)" for k in ("enc_conv5", "enc_conv7", "dec_convT2", "dec_convT3", "dec_convT4", "dec_convT5")},
             **{k: "tc_pair_kernel (cta_group::2)" for k in ("enc_conv6", "enc_conv8", "dec_dense2", "dec_convT1")},
             **{k: "tc_conv_kernel" for k in ("enc_conv4", "enc_dense")},
             "dec_dense1": "simt_conv_kernel", "latent": "latent_kernel", "enc_bn_pack": "bn_pack8_kernel"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region.

    The sampler is started before the warm-up steps (nvidia-smi needs ~0.2 s to start) and every sample carries its wall-clock
    time; stop(t0, t1) keeps the samples taken inside the timed region [t0, t1].  A short timed region (10 steps of ~9 ms) can
    fall between two samples: the caller then keeps the SAME step running untimed for a moment (t1 is extended) so that the
    clocks are still read under exactly this load, and the window is named in the result."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        # NVML (what nvidia-smi reads) polled from a thread every ~4 ms when the binding is importable: a 10-step timed region
        # of ~90 ms then holds ~20 samples; else an `nvidia-smi -lms` child as in the profiling recipe
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self._stop = False
            R = pynvml
            bits = [("hw_slowdown", R.nvmlClocksEventReasonHwSlowdown if hasattr(R, "nvmlClocksEventReasonHwSlowdown") else 0x8),
                    ("hw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonHwThermalSlowdown", 0x40)),
                    ("sw_thermal_slowdown", getattr(R, "nvmlClocksEventReasonSwThermalSlowdown", 0x20)),
                    ("sw_power_cap", getattr(R, "nvmlClocksEventReasonSwPowerCap", 0x4))]
            get_reasons = getattr(R, "nvmlDeviceGetCurrentClocksEventReasons", None) or R.nvmlDeviceGetCurrentClocksThrottleReasons
            mx = R.nvmlDeviceGetMaxClockInfo(h, R.NVML_CLOCK_SM)

            def poll():
                while not self._stop:
                    try:
                        m = get_reasons(h)
                        row = [str(R.nvmlDeviceGetClockInfo(h, R.NVML_CLOCK_SM)), str(mx), str(R.nvmlDeviceGetPowerUsage(h) / 1000.0)] + \
                              ["Active" if (m & b) else "Not Active" for _, b in bits]
                        self.rows.append((time.time(), row))
                    except Exception:
                        pass
                    time.sleep(0.004)

            self.t = threading.Thread(target=poll, daemon=True)
            self.t.start()
            self.proc = "nvml"
            return
        except Exception:
            pass
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def count(self, t0, t1):
        return sum(1 for t, _ in self.rows if t0 <= t <= t1)

    def stop(self, t0=None, t1=None, window="timed region"):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        if self.proc == "nvml":
            self._stop = True
            self.t.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if (t0 is not None and t < t0) or (t1 is not None and t > t1):
                continue
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                pw.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None,
                "samples": len(sm), "reasons": sorted(reasons), "window": window, "source": "NVML polled every ~4 ms" if self.proc == "nvml" else "nvidia-smi -lms 20"}


def synthetic_stamps_device(n, seed, device):
    """SURVEY §8d cfg 2/3 generator on the device: sky noise N(0,0.3) + a centred elliptical blob."""
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randn((n, 59, 59, 6), device=device, generator=g) * 0.3
    yy, xx = torch.meshgrid(torch.arange(59, device=device, dtype=torch.float32), torch.arange(59, device=device, dtype=torch.float32), indexing="ij")
    peak = 10 ** (torch.rand((n, 1, 1), device=device, generator=g) * 2 - 0.5)
    sx = 1.5 + 3.5 * torch.rand((n, 1, 1), device=device, generator=g)
    sy = 1.5 + 3.5 * torch.rand((n, 1, 1), device=device, generator=g)
    prof = peak * torch.exp(-0.5 * (((xx - 29) / sx) ** 2 + ((yy - 29) / sy) ** 2))
    sed = 0.3 + 0.7 * torch.rand((n, 1, 1, 6), device=device, generator=g)
    return (x + prof[..., None] * sed).contiguous()


def executed_tensor_flops(precision, stamps_per_s, pk):
    """What the tensor pipe really executes: the split precisions issue 3 MMAs per product (a_hi*w_hi + a_hi*w_lo + a_lo*w_hi),
    the fp16 tail of 'mixed' (convT6, convT7, convT8, head) 2; dec_dense1 runs on the SIMT kernel."""
    from debvader_b200.model import spec

    if precision == "fp32":
        return {}
    tail = ("dec_convT6", "dec_convT7", "dec_convT8", "dec_head")
    mult = {"bf16": 1, "bf16x3": 3, "fp16x3": 3, "mixed": 3, "fp32tc": 3}[precision]
    ex = 0
    for name, macs in spec.LAYER_MACS.items():
        if name == "dec_dense1":
            continue
        ex += 2 * macs * (2 if (precision == "mixed" and name in tail) else mult)
    tf = stamps_per_s * ex / 1e12
    return {"tflops_executed_on_tensor_pipe": round(tf, 2), "executed_frac_of_bf16_sustained": round(tf / pk["bf16_tflops_sustained"], 4),
            "executed_flop_per_stamp": ex}


def cpu_reference_rate(sample, threads=None):
    """stamps/s of the oracle port (torch-CPU restatement of reference model/model.py) on host cores."""
    from oracle import weights as ow
    from oracle.vae_torch import TorchOracle

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    o = TorchOracle(ow.make_random_weights(seed=1234), dtype=torch.float32)
    x = torch.from_numpy(ow.synthetic_stamps(min(sample, 256), seed=0))
    eps = torch.randn((x.shape[0], 32))
    o.forward(x[:32], eps[:32])  # warm-up
    done, t0 = 0, time.perf_counter()
    while done < sample:
        o.forward(x, eps)
        done += x.shape[0]
    dt = time.perf_counter() - t0
    return done / dt, threads, done, dt


def cpu_field_baseline():
    """The reference's residual field on the host (deblend/field_deblender.py:46-97: one cubic-spline
    scipy.ndimage.shift of a field-sized canvas per galaxy and band), timed through the oracle's restatement of that
    loop on a bounded sample and scaled to BASELINE cfg 4 (F=4096, 2000 sources): the cost is proportional to
    N * C * F^2, the literal run would take hours."""
    from oracle import spline_numpy as sp

    try:  # the call the reference itself makes (scipy is one of its dependencies); the numpy restatement otherwise
        from scipy.ndimage import shift as ndi_shift
        how = "scipy.ndimage.shift, the call the reference makes per galaxy and band (field_deblender.py:92-95)"
    except Exception:
        ndi_shift, how = sp.shift_cubic_constant, "numpy restatement of scipy.ndimage.shift (oracle/spline_numpy.py)"
    F, S, C, N = 1024, 59, 6, 4
    rng = np.random.default_rng(3)
    field = rng.normal(0, 0.6, (1, F, F, C))
    means = rng.random((N, S, S, C)).astype(np.float32)
    pos = rng.integers(-400, 400, size=(N, 2)).astype(np.float64)
    off = int((F - S) / 2)
    t0 = time.perf_counter()
    for k in range(N):
        for band in range(C):
            canvas = np.zeros((F, F))
            canvas[off : off + S, off : off + S] = means[k, :, :, band]
            field[0, :, :, band] -= ndi_shift(canvas, (pos[k, 0], pos[k, 1]))
    dt = time.perf_counter() - t0
    per_gal_4k = dt / N * (4096 / F) ** 2
    return {"s_per_galaxy_at_1024": dt / N, "s_per_field_4096_2000_sources_extrapolated": per_gal_4k * 2000, "cores": 1, "kind": "port",
            "sample": f"{N} galaxies x {C} bands on a {F}^2 canvas ({dt:.1f} s), scaled by (4096/{F})^2 x 2000/{N}; {how}"}


def run_reference(args, rank):
    if rank != 0:
        return
    sample = int(os.environ.get("DBV_REF_SAMPLE", "1024"))  # stamps per step (the contract test uses a small one)
    vals = []
    for i in range(args.warmup + args.steps):
        v, threads, done, dt = cpu_reference_rate(sample)
        if i >= args.warmup:
            vals.append((v, dt))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([dt for _, dt in vals]) * 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "batched deblend() of 4096 synthetic 59x59x6 stamps per GPU (BASELINE cfg 2), random-init DC2 weights",
                   "stamps_per_gpu_per_step": BATCH, "precision": "fp32 (CPU)", "sample_per_step": sample,
                   "same_config_note": f"same workload (stamp shape, architecture, weights generator) timed on a BOUNDED sample: {sample} of the {BATCH} stamps of a step per timed step; "
                                       "the metric is a rate (stamps/s), so the sample size does not enter the ratio"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} stamps per step in one batch; torch-CPU restatement of the reference model (stand-in, not TensorFlow: TF 2.13 is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def field_extras(net, device, pk, quick=False):
    """Extraction / scatter kernels (GB/s vs measured HBM) and ms per 4096^2 field (BASELINE cfg 3)."""
    from debvader_b200 import _fieldops

    F, S, C, N = 4096, 59, 6, 2000
    g = torch.Generator(device=device).manual_seed(5)
    field = (torch.randn((1, F, F, C), device=device, generator=g, dtype=torch.float32) * 0.6).double()
    rng = np.random.default_rng(5)
    centres = rng.integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2)).astype(np.float64)
    plan = _fieldops.plan_windows(centres, S, F)
    off = _fieldops.subtract_offset(F, S)
    x0, y0 = off + centres[:, 0].astype(np.int64), off + centres[:, 1].astype(np.int64)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    out = {}

    def timeit(fn, iters=5):
        fn()
        torch.cuda.synchronize()
        a, b = ev(), ev()
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / iters

    # many more stamps than one field has, so the gather is timed over >L2 of traffic: 16k stamps x 334 KB = 5.5 GB
    big_c = rng.integers(-(F // 2 - 30), F // 2 - 30, size=(16384, 2)).astype(np.float64)
    big_plan = _fieldops.plan_windows(big_c, S, F)
    t = timeit(lambda: _fieldops.extract(field, big_plan, S, C, out_dtype=torch.float64))
    out["extract_f64"] = {"GBps": 16384 * STAMP_ELTS * 16 / t / 1e6, "ms": t, "stamps": 16384, "bytes_per_stamp": STAMP_ELTS * 16}
    t = timeit(lambda: _fieldops.extract(field, big_plan, S, C, out_dtype=torch.float32))
    out["extract_f64_to_f32"] = {"GBps": 16384 * STAMP_ELTS * 12 / t / 1e6, "ms": t, "stamps": 16384, "bytes_per_stamp": STAMP_ELTS * 12}
    stamps32, _ = _fieldops.extract(field, plan, S, C, out_dtype=torch.float32)
    res = torch.empty_like(field)
    alg = N * STAMP_ELTS * 20  # stamp read + f64 field window RMW
    x0d = torch.from_numpy(x0.astype(np.int32)).to(device)  # window positions resident, like every other input of the timed region
    y0d = torch.from_numpy(y0.astype(np.int32)).to(device)
    t = timeit(lambda: _fieldops.window_axpy(field, stamps32, x0d, y0d, -1.0, out=res))
    # get_residual_field is `field.copy()` followed by the subtractions (field_deblender.py:62): the fused kernel moves the
    # whole field in+out plus the stamps once
    full = 2 * field.numel() * 8 + N * STAMP_ELTS * 4
    out["window_axpy_f64"] = {"GBps_algorithmic": alg / t / 1e6, "GBps_moved": full / t / 1e6, "ms": t, "stamps": N,
                              "note": "residual = field.copy() - stamps in one pass (binning + owner-computes kernels)"}
    for k in out.values():
        for kk in list(k):
            if kk.startswith("GBps"):
                k["frac" + kk[4:]] = round(k[kk] / pk["hbm_gbs"], 4)
    if not quick:
        # sub-pixel placement (cubic-spline ndimage.shift of field_deblender.py:92-95) of the same stamps at fractional positions
        fpos = centres + rng.uniform(-0.5, 0.5, size=centres.shape)
        E = _fieldops.spline_extent(S)
        t = timeit(lambda: _fieldops.spline_window_axpy(field, stamps32, fpos[:, 0], fpos[:, 1], -1.0), iters=3)
        tp = timeit(lambda: _fieldops.spline_place(stamps32[:512], fpos[:512, 0], fpos[:512, 1], F), iters=3)
        place_bytes = 512 * (STAMP_ELTS * 4 + E * E * C * 8)  # algorithmic: stamp in, f64 window out (the pass-X result stays in shared memory)
        out["subpixel_residual_f64"] = {"ms": t, "stamps": N, "window": E, "place_ms_per_512": tp, "place_GBps": place_bytes / tp / 1e6,
                                        "place_frac": round(place_bytes / tp / 1e6 / pk["hbm_gbs"], 4),
                                        "note": "prefilter + shift on (S+2P+2)^2 f64 windows (P=28; one warp per line, latency / issue bound), then the f64 window paste"}
        from debvader_b200.deblend_cutout.optimization import fit_positions

        # position_optimization (optimization.py:6-52) for ALL sources of the field at once: scipy's TRF path per galaxy, batched evaluations
        rb = stamps32[:, :, :, 2].double().contiguous()
        fit_positions(field, rb[:64], centres[:64])  # warm-up
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        xfit, info = fit_positions(field, rb, centres, return_info=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        out["position_fit"] = {"ms_per_galaxy": dt / N * 1e3, "galaxies": N, "ms_total": dt * 1e3, "nfev_per_galaxy": info["nfev_per_galaxy"], "rounds": info["rounds"],
                               "note": "scipy's Trust Region Reflective path restated for all galaxies of the field at once (deblend_cutout/trf_batch.py), every round of objective evaluations one batched device call (the reference runs scipy least_squares per galaxy; round 1: 9.6 ms per galaxy with scipy on the host around the device objective)"}

        def one_field():
            cut, idx = _fieldops.extract(field, plan, S, C, out_dtype=torch.float32)
            d = net(cut)
            _fieldops.center_mse(cut, d.mean().tensor, 24, 34)
            r = _fieldops.window_axpy(field, d.mean().tensor, x0, y0, -1.0, out=res)
            return _fieldops.mse(field, r)
        out["ms_per_field_kernels"] = {"value": timeit(one_field, iters=3), "field": "4096x4096x6 f64", "sources": N,
                                       "includes": "extract+cast, net, centre MSE, residual subtract, field MSE called directly on _fieldops (device resident; detection excluded)"}
        res = None
        # BASELINE metric "field deblend ms/field" through the repo's own API (SURVEY section 8d cfg 4: one deblend_field +
        # get_residual_field + field MSE, centres given): the field is a CUDA tensor, the records are lazy device-backed proxies
        from debvader_b200.deblend.field_deblender import DeblendField

        obj = DeblendField(net, field)

        def one_field_api():
            obj.deblend_field(centres)
            r = obj.get_residual_field(as_tensor=True)
            return obj.field_mse(obj.field_tensor, r)
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            t_api = timeit(one_field_api, iters=3)
        out["ms_per_field"] = {"value": t_api, "field": "4096x4096x6 f64", "sources": N,
                               "api": "DeblendField(net, field).deblend_field(centres) + get_residual_field(as_tensor=True) + field_mse",
                               "includes": "host index planning, extraction (f64 records + f32 net input), net, centre MSE + passed_cuts (one small D2H), record building, residual subtract, field MSE; detection excluded",
                               "ratio_to_kernels": round(t_api / out["ms_per_field_kernels"]["value"], 3)}
        obj = None
    # the iterative loop's in-place subtract (dbv_window_axpy_rect with in == out): only covered elements move
    work = field.clone()
    t = timeit(lambda: _fieldops.window_axpy(work, stamps32, x0d, y0d, -1.0, out=work))
    out["window_axpy_f64_inplace"] = {"GBps_algorithmic": alg / t / 1e6, "ms": t, "stamps": N, "frac_algorithmic": round(alg / t / 1e6 / pk["hbm_gbs"], 4),
                                      "note": "in-place form (out is field): read-modify-write of the covered windows + stamps only; algorithmic bytes = 417 720 B per stamp (SURVEY 8d) — overlapping windows are counted once per stamp although the kernel touches each covered element once"}
    work = None
    # BASELINE cfg 1: the packaged DC2 field (259x259x6) with its 40 catalogue centres, through the API
    try:
        g = np.load(os.path.join(ROOT, "tests", "golden", "dc2_field2.npz"))
        from debvader_b200.deblend.field_deblender import DeblendField

        f1 = torch.from_numpy(g["field"]).to(device)
        o1 = DeblendField(net, f1)

        def cfg1():
            o1.deblend_field(g["centres"])
            r = o1.get_residual_field(as_tensor=True)
            return o1.field_mse(o1.field_tensor, r)
        import contextlib
        import io

        with contextlib.redirect_stdout(io.StringIO()):
            t1 = timeit(cfg1, iters=5)
        out["cfg1_dc2_field"] = {"ms_per_field": t1, "field": "field_img_2.npy (1,259,259,6) f64 packaged with the reference", "sources": int(len(g["centres"])),
                                 "api": "DeblendField.deblend_field + get_residual_field(as_tensor=True) + field_mse"}
    except Exception as e:
        out["cfg1_dc2_field"] = {"error": repr(e)}
    # detection on the device (reference detect/detection.py:5-56 calls the CPU library sep): a 4096^2 field of noise + 2000
    # round blobs, then the iterative loop of BASELINE cfg 4 with that detector (detection -> extraction -> net -> subtract, on the device)
    try:
        from debvader_b200.detect.detection import DeviceDetector
        from debvader_b200.deblend_iterative.iterative_deblender import IterativeDeblendField

        g2 = torch.Generator(device=device).manual_seed(6)
        dfield = (torch.randn((1, F, F, C), device=device, generator=g2, dtype=torch.float32) * 0.03).double()
        yy, xx = np.mgrid[-15:16, -15:16]
        pos = rng.uniform(40, F - 40, (N, 2))
        for (px, py) in pos:
            ix, iy = int(px), int(py)
            blob = rng.uniform(0.5, 3.0) * np.exp(-((xx - (px - ix)) ** 2 + (yy - (py - iy)) ** 2) / (2 * rng.uniform(1.2, 2.5) ** 2))
            dfield[0, iy - 15 : iy + 16, ix - 15 : ix + 16, :] += torch.from_numpy(blob).to(device)[..., None]
        detector = DeviceDetector(device=device)
        t_det = timeit(lambda: detector.run(dfield), iters=5)
        t0 = time.perf_counter()
        cen = detector(dfield)
        torch.cuda.synchronize()
        t_call = (time.perf_counter() - t0) * 1e3
        out["detect"] = {"ms": t_det, "ms_call_with_centres_on_host": t_call, "objects": int(len(cen)), "sources": N, "field": "4096x4096x6 f64, r band",
                         "api": "debvader_b200.detect.detection.DeviceDetector (dbv_detect): mesh background, 7x7 matched filter, threshold, components, order, barycentres",
                         "note": "restated from the published SExtractor algorithm, bit-exact with oracle/detect_numpy.py; parity with sep itself unpinned; no multi-threshold deblending / clean pass"}
        try:  # DRAM bytes of one detection from the committed ncu pass (tools/detect_ncu_target.py, same field): the HBM roofline of the detector
            import csv

            rows = [r for r in csv.reader(open(os.path.join(ROOT, "profiles", "r02_final_detect_ncu.csv"))) if len(r) > 14 and r[0].isdigit()]
            dram = sum(float(r[14].replace(",", "")) for r in rows if r[12].startswith("dram__bytes"))
            out["detect"]["roofline"] = {"bound": "hbm", "dram_bytes_per_detection": int(dram), "achieved_GBps_moved": round(dram / t_det / 1e6, 1),
                                         "peak_GBps": pk["hbm_gbs"], "frac_moved": round(dram / t_det / 1e6 / pk["hbm_gbs"], 4),
                                         "floor_ms": round(dram / pk["hbm_gbs"] / 1e6, 3),
                                         "note": "16 small kernels, half of the bytes = the one strided pass over the 6-band field (805 MB for the r band of a 4096^2 f64 field); the rest is latency / launch bound (per-mesh serial statistics 0.30 ms)",
                                         "source": "profiles/r02_final_detect_ncu.csv"}
        except Exception:
            pass
        import contextlib
        import io

        class Bounded:  # random-init weights do not converge: at most 6 detection steps
            accepts_tensor = True

            def __init__(self):
                self.calls = 0

            def __call__(self, f):
                self.calls += 1
                return detector(f) if self.calls <= 6 else np.zeros((0, 2))

        it = IterativeDeblendField(net, dfield, detector=Bounded())
        with contextlib.redirect_stdout(io.StringIO()):
            it.iterative_deblending()  # warm-up (allocations)
            it = IterativeDeblendField(net, dfield, detector=Bounded())
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            it.iterative_deblending()
            torch.cuda.synchronize()
            t_it = (time.perf_counter() - t0) * 1e3
        out["iterative_device_detector"] = {"ms_total": t_it, "steps": len(it.nb_of_deblended_galaxies), "galaxies_per_step": [int(v) for v in it.nb_of_deblended_galaxies],
                                            "ms_per_step": t_it / max(1, len(it.nb_of_deblended_galaxies)),
                                            "api": "IterativeDeblendField(net, field, detector='device').iterative_deblending(): no field-sized transfer per step"}
        dfield = it = detector = None
    except Exception as e:
        out["detect"] = {"error": repr(e)}
    # DRAM bytes per launch of the field kernels from the committed ncu pass (same sizes as above): `traffic` next to the
    # algorithmic bytes the fractions are computed from
    try:
        ft = json.load(open(os.path.join(ROOT, "profiles", "field_traffic.json")))
        for k, ent in ft.items():
            if k in out and isinstance(ent, dict):
                out[k]["traffic"] = {"dram_bytes_per_launch": ent["dram_bytes_per_launch"], "kernel": ent["kernel"], "source": ft.get("source")}
    except Exception:
        pass
    return out


def detect_tiled_extra(device, rank, world, F=4096, N=2000):
    """SURVEY 8f-3 on `world` GPUs: TiledDeviceDetector on a 4096^2 field of noise + 2000 round blobs tiled into owner tiles + halo (one
    all-reduce of the mesh maps + one all-gather of the owned objects); ms per detection (wall clock around the synchronous call,
    max over ranks) and a check that every rank ends with the same list."""
    import torch.distributed as dist

    from debvader_b200 import parallel as par
    from debvader_b200.detect.detection import DeviceDetector, TiledDeviceDetector

    rng = np.random.default_rng(6)
    field = rng.standard_normal((1, F, F, 6), dtype=np.float32).astype(np.float64) * 0.03
    yy, xx = np.mgrid[-15:16, -15:16]
    for (px, py) in rng.uniform(40, F - 40, (N, 2)):
        ix, iy = int(px), int(py)
        blob = rng.uniform(0.5, 3.0) * np.exp(-((xx - (px - ix)) ** 2 + (yy - (py - iy)) ** 2) / (2 * rng.uniform(1.2, 2.5) ** 2))
        field[0, iy - 15 : iy + 16, ix - 15 : ix + 16, :] += blob[..., None]
    local = par.LocalField.from_full(field, rank, world, device=device)
    det = TiledDeviceDetector(device=device)
    for _ in range(2):
        c = det(local.data, local)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(5):
        c = det(local.data, local)
    torch.cuda.synchronize()
    t = torch.tensor([(time.perf_counter() - t0) / 5 * 1e3], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # every rank must hold the same list: compare a position-weighted checksum over the ranks (rank-symmetric code path; the bit-identity
    # with the single-GPU list is checked by tools/detect_tiled_nccl.py, profiles/r02_detect_tiled_2gpu.json, and at world == 1 here)
    w = np.arange(1, len(c) + 1, dtype=np.float64)
    chk = torch.tensor([float(len(c)), float((c[:, 0] * w).sum()), float((c[:, 1] * w * 3.0).sum())], device=device, dtype=torch.float64)
    lo, hi = chk.clone(), chk.clone()
    if world > 1:
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same_everywhere = bool(torch.equal(lo, hi))
    identical_single = None
    if world == 1:
        identical_single = bool(np.array_equal(c, DeviceDetector(device=device)(local.data)))
    return {"ms_per_detection": float(t.item()), "objects": int(len(c)), "same_list_on_every_rank": same_everywhere, "identical_to_plain_detector": identical_single,
            "assembled_field_fallbacks": det.fallbacks, "field_share_per_rank": round(local.data.numel() / (F * F * 6), 4),
            "collectives": "one all_reduce(MAX) of the (2, 64, 64) mesh maps + two all_gathers (counts, owned objects) per detection",
            "timing": "wall clock around the synchronous call (it returns host centres), max over ranks"}


def field_tiled_extra(net, device, rank, world, quick=False):
    """BASELINE cfg 4 on `world` GPUs, ALL ranks taking part: a 4096^2 x 6 f64 field with 2000 sources tiled into owner
    tiles + 30-px halo (each rank uploads only its local region), one pass = DeblendField(tiled=True).deblend_field +
    get_residual_field(as_tensor=True) + field_mse (one NCCL all_to_all of overlapping stamps + one all-reduce of a double).
    Timed with CUDA events, max over ranks.  The regions are compared bit for bit with the single-GPU residual, computed
    by rank 0 alone on the full field and sent region by region."""
    import contextlib
    import io

    import torch.distributed as dist

    from debvader_b200 import parallel as par
    from debvader_b200.deblend.field_deblender import DeblendField

    F, S, C, N = (1025, 59, 6, 300) if quick else (4096, 59, 6, 2000)
    field = np.random.default_rng(5).standard_normal((1, F, F, C), dtype=np.float32).astype(np.float64) * 0.6  # same array in every process
    centres = np.random.default_rng(6).integers(-(F // 2 - 30), F // 2 - 30, size=(N, 2)).astype(np.float64)
    was = net.sample
    net.sample = False  # z = loc: deterministic pass, so that tiles can be compared bit for bit
    try:
        obj = DeblendField(net, field, tiled=True)
        keep = {}

        def one():
            obj.deblend_field(centres)
            keep["res"] = obj.get_residual_field(as_tensor=True)
            keep["mse"] = obj.field_mse(obj.field_tensor, keep["res"])

        with contextlib.redirect_stdout(io.StringIO()):
            for _ in range(3):  # the steady state alternates between two result buffers (the caller still holds the previous
                one()           # residual while the next one is made): a first cudaMalloc of 805 MB costs ~100 ms; NCCL's
                                # point-to-point channels also settle over the first passes
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            import gc

            gc_log = []

            def _gc_cb(phase, info, _t=[0.0]):
                if phase == "start":
                    _t[0] = time.perf_counter()
                elif info.get("generation", 0) >= 2:
                    gc_log.append(round((time.perf_counter() - _t[0]) * 1e3, 2))

            gc.callbacks.append(_gc_cb)
            iters = 8
            evs = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
            evs[0].record()
            for i in range(iters):
                one()
                evs[i + 1].record()
            torch.cuda.synchronize()
            gc.callbacks.remove(_gc_cb)
            a, b = evs[0], evs[-1]
            each = [round(evs[i].elapsed_time(evs[i + 1]), 3) for i in range(iters)]
        ms = torch.tensor([a.elapsed_time(b) / iters], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        loc = obj._local
        share = loc.nbytes() / (F * F * C * 8)
        same, mse_single = True, None
        if rank == 0:
            with contextlib.redirect_stdout(io.StringIO()):
                single = DeblendField(net, torch.from_numpy(field).to(device))
                single.deblend_field(centres)
                full = single.get_residual_field(as_tensor=True)
                mse_single = single.field_mse(single.field_tensor, full)
            for r, (a0, a1, b0, b1) in enumerate(par.region_bounds(F, world)):
                part = full[:, a0:a1, b0:b1].contiguous()
                if r == 0:
                    same = bool(torch.equal(part, keep["res"]))
                else:
                    dist.send(part, dst=r)
            del full, single
        else:
            truth = torch.empty_like(keep["res"])
            dist.recv(truth, src=0)
            same = bool(torch.equal(truth, keep["res"]))
        flag = torch.tensor([1 if same else 0], device=device)
        if world > 1:
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        mse_tiled = keep["mse"]
        del obj
        keep.clear()
        torch.cuda.empty_cache()
        try:
            det_tiled = detect_tiled_extra(device, rank, world, F, N)
        except Exception as e:  # never breaks the contract line
            det_tiled = {"error": repr(e)}
        return {"ms_per_field": float(ms.item()), "n_gpus": world, "field": f"{F}x{F}x{C} f64", "sources": N, "detect_tiled": det_tiled,
                "tiles": list(par.tile_grid(world)), "halo_px": par.HALO, "field_share_per_rank": round(share, 4),
                "bit_identical_to_single_gpu": bool(flag.item()), "mse_tiled": mse_tiled, "mse_single": mse_single,
                "ms_each_pass_rank0": each, "full_gc_pauses_ms_rank0": gc_log,
                "api": "DeblendField(net, field, tiled=True).deblend_field + get_residual_field(as_tensor=True) + field_mse",
                "collectives": "one all_to_all_single of overlapping stamps + one all_reduce of a double per pass", "timing": "CUDA events, max over ranks"}
    finally:
        net.sample = was


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="mixed", choices=["mixed", "bf16x3", "fp16x3", "bf16", "fp32", "fp32tc"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--chunk", type=int, default=0)
    ap.add_argument("--no-extras", action="store_true", help="skip cpu_baseline / field / alternative-precision extras")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist

    from debvader_b200 import _ffi
    from debvader_b200.deblend_cutout.deblender import deblend
    from debvader_b200.model import spec
    from debvader_b200.model.model import load_deblender

    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU path)"
    torch.cuda.set_device(local)
    numa = "unchanged"
    if world > 1:
        # one process per GPU: run it (and first-touch its pinned host buffers) on the CPUs / NUMA node next to its GPU, so that
        # the host<->device copies of the end-to-end path do not all cross the socket interconnect
        try:
            import pynvml

            pynvml.nvmlInit()
            pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
            numa = f"cpu affinity set to GPU {local}'s local CPUs ({len(os.sched_getaffinity(0))} of {os.cpu_count()})"
        except Exception as e:  # the measurement still stands, only slower copies
            numa = f"not set ({type(e).__name__})"
    device = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
    pk = peaks()
    B = args.batch
    net = load_deblender(*CFG, weights="random:1234", precision=args.precision, chunk=args.chunk, seed=rank)
    x = synthetic_stamps_device(B, 1000 + rank, device)
    mean = torch.empty((B, 59, 59, 6), device=device)
    std = torch.empty_like(mean)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)  # nvidia-smi start-up
    torch.cuda.synchronize()
    t_warm0 = time.time()
    for _ in range(args.warmup):
        net.deblend_into(x, mean, std)
    net.set_profiling(True)  # per-layer CUDA events on the launching stream, inside the timed region
    barrier()
    t_wall0 = time.time()
    l0 = _ffi.lib().dbv_global_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.profiler.start()  # `ncu --profile-from-start off` then sees only the timed region (not the plan autotuner)
    e0.record()
    for _ in range(args.steps):
        net.deblend_into(x, mean, std)
    e1.record()
    barrier()
    torch.cuda.profiler.stop()
    launches = int(_ffi.lib().dbv_global_launch_count() - l0)
    t_wall1 = time.time()
    clk = None
    layers_timed = None
    if rank == 0:
        window = "timed region"
        if clocks.proc is not None and clocks.count(t_wall0, t_wall1) < 3 and clocks.count(t_warm0, t_wall1) >= 3:
            t_wall0, window = t_warm0, "warm-up steps + timed region (the same step, back to back; the timed region alone is shorter than three sampling periods)"
        elif clocks.proc is not None and clocks.count(t_wall0, t_wall1) < 3:
            # too short for the sampler: keep the identical step running (untimed, not profiled) until it has read the clocks
            layers_timed = net.layer_times()
            net.set_profiling(False)
            t_end = time.time() + 0.6
            while time.time() < t_end:
                net.deblend_into(x, mean, std)
                torch.cuda.synchronize()
            t_wall1 = time.time()
            window = "timed region + 0.6 s of the same step repeated right after it (the timed region is shorter than the sampling period)"
        clk = clocks.stop(t_wall0, t_wall1, window)
    ms = torch.tensor([e0.elapsed_time(e1)], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    value = world * B * args.steps / (ms_total / 1e3)
    layers = layers_timed if layers_timed is not None else net.layer_times()  # per-layer CUDA-event times, averaged over the K timed steps
    net.set_profiling(False)

    # ---- e2e through the public API with host buffers ------------------------------------------------
    x_host = torch.empty((B, 59, 59, 6), dtype=torch.float32, pin_memory=True)
    x_host.copy_(x)
    xh = x_host.numpy()
    for _ in range(2):
        m, d = deblend(net, xh)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        m, d = deblend(net, xh)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e = world * B * e2e_steps / float(t_e2e.item())

    # ---- e2e with the reference's natural input: a pageable float64 ndarray (deblend_cutout/deblender.py:18 casts it) --------
    e2e_f64 = e2e_f32p = None
    if not args.no_extras:
        xh64 = xh.astype(np.float64)  # pageable
        m, d = deblend(net, xh64)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            m, d = deblend(net, xh64)
        torch.cuda.synchronize()
        t64 = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t64, op=dist.ReduceOp.MAX)
        e2e_f64 = world * B * 3 / float(t64.item())
        del xh64
        xh32 = np.array(xh, copy=True)  # pageable float32: an ordinary numpy array
        m, d = deblend(net, xh32)
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            m, d = deblend(net, xh32)
        torch.cuda.synchronize()
        t32 = torch.tensor([time.perf_counter() - t0], device=device, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t32, op=dist.ReduceOp.MAX)
        e2e_f32p = world * B * 3 / float(t32.item())
        del xh32

    # ---- BASELINE cfg 4: the tiled field pass, every rank taking part ------------------------------------------------
    field_tiled = None
    if not args.no_extras:
        try:
            field_tiled = field_tiled_extra(net, device, rank, world)
        except Exception as e:  # extras never break the contract line
            field_tiled = {"error": repr(e)}

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ---------------------------------------------------------------
    # MEASURED_PEAKS carries a burst figure (kernel timed alone / short region) and a sustained one (inside a long step):
    # the timed region is K steps of ~7 ms, i.e. burst conditions unless it lasts >= 1 s; both fractions are reported
    burst = ms_total < 1000.0
    peak_tf = pk["bf16_tflops"] if burst else pk["bf16_tflops_sustained"]
    lay = []
    for name, lms in layers:
        macs = spec.LAYER_MACS.get(name, 0)
        tf = 2 * macs * B / (lms / 1e3) / 1e12 if lms > 0 else 0.0
        lay.append({"layer": name, "ms": round(lms, 4), "tflops": round(tf, 2), "frac": round(tf / peak_tf, 4)})
    tc_lay = [l for l in lay if spec.LAYER_MACS.get(l["layer"], 0) > 0]
    slowest = max(tc_lay, key=lambda l: l["ms"]) if tc_lay else {"layer": None, "tflops": 0.0, "ms": 0.0}
    sum_ms = sum(l["ms"] for l in lay) or 1.0
    net_tf = value / world * spec.FLOP_PER_STAMP / 1e12
    # The dominant KERNEL is the __global__ function with the largest share of the step, named as the ncu launch list of the same
    # command names it (profiles/*_launches.csv aggregates by function: one instantiation runs several layers).  Its achieved rate
    # = algorithmic FLOPs of all its launches in a step / the CUDA-event time of those launches (i.e. per average launch).
    groups = {}
    for l in lay:
        try:
            kn = net.layer_kernel(l["layer"])
        except Exception:
            kn = KERNEL_OF.get(l["layer"], "?")
        l["kernel"] = kn
        g = groups.setdefault(kn, {"kernel": kn, "layers": [], "ms": 0.0, "flop": 0.0})
        g["layers"].append(l["layer"])
        g["ms"] += l["ms"]
        g["flop"] += 2.0 * spec.LAYER_MACS.get(l["layer"], 0) * B
    for g in groups.values():
        g["tflops"] = round(g["flop"] / (g["ms"] / 1e3) / 1e12, 2) if g["ms"] > 0 else 0.0
        g["share_of_step"] = round(g["ms"] / sum_ms, 4)
        g["frac_of_burst"] = round(g["tflops"] / pk["bf16_tflops"], 4)
        g["ms"] = round(g["ms"], 4)
    kernels = sorted(({k: v for k, v in g.items() if k != "flop"} for g in groups.values()), key=lambda g: -g["ms"])
    tc_groups = [g for g in groups.values() if g["flop"] > 0]
    top = max(tc_groups, key=lambda g: g["ms"]) if tc_groups else {"kernel": None, "layers": [], "tflops": 0.0, "ms": 0.0, "flop": 0.0}
    traffic, traffic_detail = None, None
    try:  # dram__bytes_read.sum + dram__bytes_write.sum of this kernel's launches, from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ents = [tj.get(args.precision, {}).get(n) for n in top["layers"]]
        if ents and all(ents):
            # per average launch of THIS run (B stamps), scaled from the captured launches (traffic is linear in the stamps: no reuse across stamps)
            per_launch = [e["dram_bytes"] / e["stamps"] * B for e in ents]
            traffic = int(sum(per_launch) / len(per_launch))
            traffic_detail = {"per_layer_bytes_per_launch": {n: int(v) for n, v in zip(top["layers"], per_launch)}, "stamps_per_launch": B,
                              "algorithmic_flops_per_average_launch": int(top["flop"] / len(top["layers"])), "source": tj.get("source")}
    except Exception:
        pass
    n_l = max(1, len(top["layers"]))
    n_chunks = max(1, -(-B // (args.chunk or 4096)))
    roofline = {"bound": "tensor", "kernel": f"{top['kernel']} (tcgen05; runs {', '.join(top['layers'])})",
                "achieved": top["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                "frac": round(top["tflops"] / peak_tf, 4), "traffic": traffic, "traffic_detail": traffic_detail, "share_of_step": round(top["ms"] / sum_ms, 4),
                "launches_per_step": n_l * n_chunks, "avg_launch_ms": round(top["ms"] / n_l / n_chunks, 4),
                "frac_of_burst": round(top["tflops"] / pk["bf16_tflops"], 4), "frac_of_sustained": round(top["tflops"] / pk["bf16_tflops_sustained"], 4),
                "peak_source": pk["source"] + (" bf16 burst (timed region %.2f s < 1 s)" % (ms_total / 1e3) if burst else " bf16 sustained (timed region %.2f s)" % (ms_total / 1e3)),
                "flops": "algorithmic (nominal 2*MACs of the layers; the hi/lo split precisions execute 3x that on the tensor pipe, 2x in the fp16 tail of 'mixed')",
                "dominant_by": "largest share of the step among the __global__ functions, as the ncu launch list groups them (profiles/r02_final_launches.csv)",
                "slowest_layer": {"layer": slowest["layer"], "kernel": slowest.get("kernel"), "ms": slowest["ms"], "tflops": slowest["tflops"],
                                  "frac_of_burst": round(slowest["tflops"] / pk["bf16_tflops"], 4), "share_of_step": round(slowest["ms"] / sum_ms, 4),
                                  "note": "enc_conv1 has K = 54 (6 input bands): 6 MFLOP per stamp against 0.5 MB of activations written as 16-bit hi/lo pairs — it is bound by its epilogue's stores (ncu: LSU 74 %, DRAM 3.1 TB/s), not by the tensor pipe"},
                "kernels": kernels}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"fp32": "f32", "fp16x3": "fp16", "fp32tc": "fp16 hi/lo, fp32-promoted partial sums", "mixed": "bf16+fp16"}.get(args.precision, "bf16"), "data": "synthetic",
        "config": {"workload": "batched deblend() of 4096 synthetic 59x59x6 stamps per GPU (BASELINE cfg 2), random-init DC2 weights",
                   "stamps_per_gpu_per_step": B, "precision": args.precision, "l2": "step input 342 MB > 126 MB L2 (inputs larger than L2)",
                   "parallelism": f"dp{world} (stamps sharded, no data-path collective)"},
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": B * STAMP_ELTS * 4, "d2h_bytes_per_step": B * STAMP_ELTS * 4,
                "api": "debvader_b200.deblend_cutout.deblender.deblend(net, host ndarray) -> dbv_deblend_host", "timing": "wall clock, max over ranks",
                "note": "returns the mean ndarray (device->host copy inside the timed region) and the distribution object, whose stddev stays on the device until a caller asks for it; input = pinned float32. Copy ceiling of the pool's boxes (tools/pcie_probe.py, profiles/r02_pcie_probe_*gpu.log): 551-593 k stamps/s on one GPU (55 GB/s each way), 765 k stamps/s in total on 8 GPUs (23 GB/s H2D, 11.7 GB/s D2H per rank when eight ranks copy at once): at N = 8 this figure IS the box's ceiling, not a code limit",
                "pageable_f64_input": {"value": e2e_f64, "unit": UNIT, "h2d_bytes_per_step": B * STAMP_ELTS * 4,
                                       "note": "the reference's natural input: a pageable float64 ndarray (deblender.py:18 casts it); host threads of the library convert it to float32 into pinned staging memory piece by piece (csrc/host_stage.cu), the H2D copy is then a DMA transfer of half the bytes; round 1 / early round 2: 62-66 k stamps/s with the driver staging 684 MB on one thread"},
                "pageable_f32_input": {"value": e2e_f32p, "unit": UNIT, "h2d_bytes_per_step": B * STAMP_ELTS * 4,
                                       "note": "an ordinary (pageable) float32 ndarray, staged into pinned memory by the library's host threads"},
                "host_affinity": numa},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": roofline,
        "network": {"tflops_algorithmic": round(net_tf, 2), "frac_of_bf16_sustained": round(net_tf / pk["bf16_tflops_sustained"], 4),
                    "frac_of_bf16_burst": round(net_tf / pk["bf16_tflops"], 4), "flop_per_stamp": spec.FLOP_PER_STAMP,
                    **executed_tensor_flops(args.precision, value / world, pk)},
        "layers": lay,
        "field_tiled": field_tiled,
    }
    # extras (other precisions, field kernels, CPU baselines) belong to the N=1 line only: at N>1 the other ranks would sit in
    # the final barrier (spinning on their host threads) while rank 0 runs them
    if not args.no_extras and world == 1:
        try:
            alt = {}
            for prec in [p for p in ("bf16", "bf16x3", "fp16x3", "mixed", "fp32tc", "fp32") if p != args.precision]:
                n2 = load_deblender(*CFG, weights="random:1234", precision=prec, chunk=args.chunk)
                reps = 2 if prec == "fp32" else 5
                for _ in range(1 if prec == "fp32" else 3):
                    n2.deblend_into(x, mean, std)
                torch.cuda.synchronize()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                for _ in range(reps):
                    n2.deblend_into(x, mean, std)
                b.record()
                torch.cuda.synchronize()
                v = B * reps / (a.elapsed_time(b) / 1e3)
                alt[prec] = {"value": v, "unit": UNIT, "n_gpus": 1,
                             "note": {"bf16": "single-pass bf16: ~1e-2 of peak flux, does NOT meet the 1e-3 tolerance", "mixed": "meets 1e-3 (measured ~5e-4 of peak flux)",
                                      "fp32": "fp32 SIMT tier (FFMA2 implicit GEMM through shared memory): meets 1e-5 (measured ~2e-6 of peak flux)",
                                      "fp32tc": "the 1e-5 tier on tcgen05: fp16 hi/lo operands, accumulation chains cut every <= 128 values of K and promoted into fp32 registers (add.rn)"}.get(prec, "meets 1e-3 (measured ~5e-5 of peak flux)")}
                if prec == "fp32":
                    alt[prec]["tflops_fp32"] = round(v * spec.FLOP_PER_STAMP / 1e12, 2)
                n2.close()
            line["alt_precision"] = alt
        except Exception as e:  # extras never break the contract line
            line["alt_precision"] = {"error": repr(e)}
        try:
            line["field"] = field_extras(net, device, pk)
        except Exception as e:
            line["field"] = {"error": repr(e)}
        try:
            v, threads, done, dt = cpu_reference_rate(8192)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
                                    "sample": f"{done} stamps in batches of 256 ({dt:.1f} s); torch-CPU restatement of the reference model (stand-in, not TensorFlow)"}
        except Exception as e:
            line["cpu_baseline"] = {"error": repr(e)}
        try:
            if isinstance(line.get("field"), dict) and "ms_per_field" in line["field"]:
                line["field"]["cpu_baseline"] = cpu_field_baseline()
        except Exception as e:
            line["field"]["cpu_baseline"] = {"error": repr(e)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
