"""B200-native drop-in for reference ``debvader.model.model`` (model/model.py).

``load_deblender`` / ``create_model_vae`` keep the reference signatures
(model/model.py:164-172, 221-229) and return callables with the same surface as
the Keras models (``net(x)`` -> distribution with ``.mean()/.stddev()/.sample()/
.log_prob()``, ``encoder(x)`` -> (B,560), ``decoder(z)``, ``z(x)``), but every
call runs hand-written sm_100a kernels through the C-ABI in
``include/debvader_b200.h``.  Only the DC2 architecture is compiled in; any
other configuration raises (there is no fallback path).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np
import torch

from .. import _ffi
from .._dist import MVNTriLOutput, NormalOutput, Value
from . import ckpt, spec

S, NB, LAT, NPAR = 59, 6, 32, 560
# "mixed" = bf16 hi/lo split everywhere except single-plane fp16 activations into the four large-image decoder layers:
# ~5e-4 of peak flux (north_star tolerance for the tensor-core path: 1e-3), fp16 range (+-65504) on those activations.
# "bf16x3" (~5e-5, full fp32 range) and "fp32" (SIMT, <= 1e-5) are the tighter choices.
DEFAULT_PRECISION = os.environ.get("DEBVADER_B200_PRECISION", "mixed")


def _as_device_f32(x, device):
    """tf.cast(images, tf.float32) (deblend_cutout/deblender.py:18) onto `device`."""
    if isinstance(x, Value):
        x = x.tensor
    if not isinstance(x, torch.Tensor):
        x = torch.from_numpy(np.ascontiguousarray(x))
    x = x.to(device, non_blocking=True)
    if x.dtype != torch.float32:
        x = x.float()
    return x.contiguous()


class Deblender:
    """``net`` of create_model_vae (model/model.py:216): encoder -> MultivariateNormalTriL -> decoder."""

    def __init__(self, weights: dict, precision: str | None = None, device: int | None = None, chunk: int = 0, seed: int = 0):
        precision = precision or DEFAULT_PRECISION
        if precision not in _ffi.PREC:
            raise ValueError(f"precision must be one of {sorted(_ffi.PREC)}, got {precision!r}")
        lib = _ffi.lib()
        if not torch.cuda.is_available():
            raise RuntimeError("debvader_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        self.precision = precision
        self._seed = int(seed)
        self._calls = 0
        self.sample = True  # default of net(x): z is SAMPLED like the reference; set False for z = loc (deterministic passes)
        self._ctx = C.c_void_p()
        _ffi.check(lib.dbv_create(C.byref(self._ctx), self.device_index, _ffi.PREC[precision], int(chunk)))
        table = dict(spec.tensor_table())
        missing = [k for k in table if k not in weights]
        if missing:
            raise KeyError(f"weights are missing {len(missing)} tensors, e.g. {missing[0]}")
        for key in table:
            a = np.ascontiguousarray(np.asarray(weights[key], dtype=np.float32))
            shp = (C.c_int64 * a.ndim)(*a.shape)
            _ffi.check(lib.dbv_set_weights(self._ctx, key.encode(), a.ctypes.data_as(C.c_void_p), shp, a.ndim))
        _ffi.check(lib.dbv_finalize_weights(self._ctx))
        self.trainable = False

    # ---- lifecycle ---------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None) is not None and self._ctx.value:
            _ffi.lib().dbv_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def seed(self, seed: int):
        self._seed, self._calls = int(seed), 0

    def _next_seed(self, seed):
        if seed is not None:
            return int(seed) & (2**64 - 1)
        self._calls += 1
        return (self._seed * 0x9E3779B97F4A7C15 + self._calls) & (2**64 - 1)

    def fp16_overflow(self, reset: bool = False) -> bool:
        """True if an activation of the fp16 tail of precision="mixed" has saturated at +-65504 in a call that has
        completed (sticky).  Such a result is outside the 1e-3 tolerance: use precision="bf16x3" (fp32 range)."""
        return bool(_ffi.lib().dbv_fp16_overflow(self._ctx, int(bool(reset))))

    def _raise_on_overflow(self):
        if self.precision == "mixed" and self.fp16_overflow(reset=True):
            raise FloatingPointError(
                'precision="mixed" keeps the inputs of the four large-image decoder layers in fp16 and one of them left the fp16 range '
                "(+-65504) in an earlier call: those results are not within the 1e-3 tolerance. Use precision=\"bf16x3\" (fp32 range, ~16 % slower)."
            )

    @property
    def launches(self) -> int:
        return int(_ffi.lib().dbv_launch_count(self._ctx))

    # ---- stages on device tensors ---------------------------------------------------------------
    def encode(self, x) -> torch.Tensor:
        """encoder(x): model/model.py:61-100.  (B,59,59,6) -> (B,560) fp32 on the device."""
        x = self._check_x(_as_device_f32(x, self.device))
        with torch.cuda.device(self.device):
            out = torch.empty((x.shape[0], NPAR), device=self.device, dtype=torch.float32)
            _ffi.check(_ffi.lib().dbv_encode(self._ctx, _ffi.ptr(x), x.shape[0], _ffi.ptr(out), _ffi.stream_ptr()))
        return out

    def latent(self, params, eps=None, sample=True, seed=None):
        """MultivariateNormalTriL(32): model/model.py:43-58, 211-214 -> (z, loc, stddev)."""
        params = _as_device_f32(params, self.device)
        B = params.shape[0]
        if eps is not None:
            eps = _as_device_f32(eps, self.device)
            if tuple(eps.shape) != (B, LAT):
                raise ValueError(f"eps must have shape {(B, LAT)}, got {tuple(eps.shape)}")
        with torch.cuda.device(self.device):
            z = torch.empty((B, LAT), device=self.device, dtype=torch.float32)
            loc = torch.empty_like(z)
            std = torch.empty_like(z)
            _ffi.check(
                _ffi.lib().dbv_latent(self._ctx, _ffi.ptr(params), _ffi.ptr(eps), self._next_seed(seed), int(bool(sample)), 0, B,
                                      _ffi.ptr(z), _ffi.ptr(loc), _ffi.ptr(std), _ffi.stream_ptr())
            )
        return z, loc, std

    def decode(self, z) -> NormalOutput:
        """decoder(z): model/model.py:103-161."""
        z = _as_device_f32(z, self.device)
        if z.ndim != 2 or z.shape[1] != LAT:
            raise ValueError(f"z must have shape (B,{LAT}), got {tuple(z.shape)}")
        B = z.shape[0]
        with torch.cuda.device(self.device):
            mean = torch.empty((B, S, S, NB), device=self.device, dtype=torch.float32)
            std = torch.empty_like(mean)
            _ffi.check(_ffi.lib().dbv_decode(self._ctx, _ffi.ptr(z), B, _ffi.ptr(mean), _ffi.ptr(std), _ffi.stream_ptr()))
        return NormalOutput(mean, std)

    def __call__(self, x, eps=None, sample=None, seed=None, return_z=False):
        """net(x) (deblend_cutout/deblender.py:18) on device-resident data.

        By default z is *sampled* like the reference (the TFP layer's convert_to_tensor_fn is
        Distribution.sample); pass ``eps=`` for a given draw or ``sample=False`` for z = loc
        (``net.sample = False`` changes the default of the instance)."""
        sample = self.sample if sample is None else sample
        self._raise_on_overflow()  # of an EARLIER call (this one is asynchronous); synchronous callers use fp16_overflow()
        x = self._check_x(_as_device_f32(x, self.device))
        B = x.shape[0]
        if eps is not None:
            eps = _as_device_f32(eps, self.device)
            if tuple(eps.shape) != (B, LAT):
                raise ValueError(f"eps must have shape {(B, LAT)}, got {tuple(eps.shape)}")
        with torch.cuda.device(self.device):
            mean = torch.empty((B, S, S, NB), device=self.device, dtype=torch.float32)
            std = torch.empty_like(mean)
            z = torch.empty((B, LAT), device=self.device, dtype=torch.float32) if return_z else None
            _ffi.check(
                _ffi.lib().dbv_deblend(self._ctx, _ffi.ptr(x), B, _ffi.ptr(eps), self._next_seed(seed), int(bool(sample)),
                                       _ffi.ptr(mean), _ffi.ptr(std), _ffi.ptr(z), _ffi.stream_ptr())
            )
        out = NormalOutput(mean, std)
        return (out, z) if return_z else out

    def epistemic_std(self, x, n: int = 100, seed=None) -> torch.Tensor:
        """Per-pixel std of the predicted mean over `n` latent draws per stamp — the epistemic uncertainty of
        deblend/field_deblender.py:303-316 (np.std(deblend(net, [cutout] * 100)[0], axis=0)), batched: the encoder is
        deterministic, so it runs ONCE per stamp; only the latent sampling and the decoder run n times.
        (B,59,59,6) -> (B,59,59,6) float64 CUDA tensor (ddof = 0 like np.std)."""
        params = self.encode(x)
        B = params.shape[0]
        out = torch.empty((B, S, S, NB), device=self.device, dtype=torch.float64)
        per = max(1, 4096 // int(n))  # stamps per decoder call (n draws each)
        for s in range(0, B, per):
            e = min(B, s + per)
            p = params[s:e].repeat_interleave(int(n), dim=0).contiguous()
            z, _, _ = self.latent(p, sample=True, seed=None if seed is None else int(seed) + s)
            m = self.decode(z).mean().tensor.view(e - s, int(n), S, S, NB)
            out[s:e] = m.double().std(dim=1, unbiased=False)
        return out

    def deblend_into(self, x, mean, stddev=None, z=None, eps=None, sample=True, seed=None):
        """net(x) into caller-provided CUDA tensors (no allocation; what bench.py times)."""
        B = self._check_x(x).shape[0]
        for name, t, shape in (("x", x, (B, S, S, NB)), ("mean", mean, (B, S, S, NB)), ("stddev", stddev, (B, S, S, NB)), ("z", z, (B, LAT)), ("eps", eps, (B, LAT))):
            if t is None:
                continue
            if not (isinstance(t, torch.Tensor) and t.device == self.device and t.dtype == torch.float32 and t.is_contiguous() and tuple(t.shape) == shape):
                raise ValueError(f"{name} must be a contiguous float32 tensor of shape {shape} on {self.device}")
        _ffi.check(
            _ffi.lib().dbv_deblend(self._ctx, _ffi.ptr(x), x.shape[0], _ffi.ptr(eps), self._next_seed(seed), int(bool(sample)),
                                   _ffi.ptr(mean), _ffi.ptr(stddev), _ffi.ptr(z), _ffi.stream_ptr())
        )

    # ---- host buffers: the end-to-end call ---------------------------------------------------------
    def deblend_host(self, images, eps=None, sample=None, seed=None, want_stddev=True, out_mean=None, out_stddev=None, resident=False):
        """Host ndarray in, host ndarrays out, H2D / compute / D2H pipelined inside the C-ABI
        (dbv_deblend_host).  float64 input is cast on the device.  Returns (mean, stddev|None).

        resident=True keeps both outputs on the device as well and copies only the MEAN back
        (``deblend()`` returns the mean ndarray plus a distribution whose stddev is fetched on demand);
        it returns (mean ndarray, mean CUDA tensor, stddev CUDA tensor)."""
        sample = self.sample if sample is None else sample
        a = images if isinstance(images, np.ndarray) else np.asarray(images)
        if a.dtype not in (np.float32, np.float64):
            a = a.astype(np.float32)
        a = np.ascontiguousarray(a)
        self._check_x(a)
        B = a.shape[0]
        # outputs live in pinned host memory (torch's caching host allocator recycles the blocks), so the
        # D2H copies of the pipeline are asynchronous DMA transfers
        new = lambda: torch.empty((B, S, S, NB), dtype=torch.float32, pin_memory=True).numpy()
        for name, arr in (("out_mean", out_mean), ("out_stddev", out_stddev)):
            if arr is not None and not (isinstance(arr, np.ndarray) and arr.dtype == np.float32 and arr.flags.c_contiguous and arr.shape == (B, S, S, NB)
                                        and arr.flags.writeable):
                raise ValueError(f"{name} must be a writeable C-contiguous float32 ndarray of shape {(B, S, S, NB)}")
        mean = out_mean if out_mean is not None else new()
        std = (out_stddev if out_stddev is not None else new()) if (want_stddev and not resident) else None
        mean_dev = std_dev = None
        if resident:
            with torch.cuda.device(self.device):
                mean_dev = torch.empty((B, S, S, NB), device=self.device, dtype=torch.float32)
                std_dev = torch.empty_like(mean_dev)
                # the library computes on its own streams (ordered after this ctx's earlier device-pointer calls by an
                # event inside the library): make sure nothing queued on torch's stream still uses these fresh blocks
                torch.cuda.current_stream().synchronize()
        e = None
        if eps is not None:
            e = np.ascontiguousarray(np.asarray(eps, dtype=np.float32))
            if e.shape != (B, LAT):
                raise ValueError(f"eps must have shape {(B, LAT)}, got {e.shape}")
        vp = lambda arr: None if arr is None else arr.ctypes.data_as(C.c_void_p)
        _ffi.check(
            _ffi.lib().dbv_deblend_host(self._ctx, vp(a), _ffi.F64 if a.dtype == np.float64 else _ffi.F32, B, vp(e),
                                        self._next_seed(seed), int(bool(sample)), vp(mean), vp(std), None,
                                        _ffi.ptr(mean_dev), _ffi.ptr(std_dev))
        )
        self._raise_on_overflow()  # dbv_deblend_host is synchronous: this call's own flag
        if resident:
            return mean, mean_dev, std_dev
        return mean, std

    # ---- diagnostics ------------------------------------------------------------------------------
    def set_profiling(self, on: bool):
        _ffi.check(_ffi.lib().dbv_set_profiling(self._ctx, int(bool(on))))

    def layer_times(self):
        """[(layer name, ms)] per call, averaged over the calls made since profiling was switched on (CUDA events)."""
        n = 64
        ms = (C.c_float * n)()
        names = C.create_string_buffer(32 * n)
        torch.cuda.synchronize(self.device)
        k = _ffi.check(_ffi.lib().dbv_layer_times(self._ctx, n, ms, names))
        return [(names.raw[32 * i : 32 * i + 32].split(b"\0", 1)[0].decode(), float(ms[i])) for i in range(k)]

    def layer_kernel(self, name: str) -> str:
        """the __global__ function (as the ncu launch list names it) that runs `name` under this precision and tuned plan"""
        buf = C.create_string_buffer(64)
        _ffi.check(_ffi.lib().dbv_layer_kernel(self._ctx, name.encode(), buf, 64))
        return buf.value.decode()

    def debug_activation(self, name: str, B: int, shape):
        out = torch.empty((B,) + tuple(shape), device=self.device, dtype=torch.float32)
        with torch.cuda.device(self.device):
            _ffi.check(_ffi.lib().dbv_debug_activation(self._ctx, name.encode(), B, _ffi.ptr(out), _ffi.stream_ptr()))
        return out

    @staticmethod
    def _check_x(x):
        if x.ndim != 4 or tuple(x.shape[1:]) != (S, S, NB):
            raise ValueError(f"images must have shape (B,{S},{S},{NB}) (the DC2 deblender), got {tuple(x.shape)}")
        return x


class EncoderModel:
    """``encoder`` of create_model_vae (model/model.py:182-189): x -> (B,560)."""

    def __init__(self, net: Deblender):
        self.net = net

    def __call__(self, x):
        return Value(self.net.encode(x))


class DecoderModel:
    """``decoder`` of create_model_vae (model/model.py:191-199): z -> Normal(loc, scale)."""

    def __init__(self, net: Deblender):
        self.net = net
        self.trainable = False

    def __call__(self, z):
        if isinstance(z, MVNTriLOutput):
            z = z.sample().tensor
        return self.net.decode(z)


class LatentModel:
    """``Model(inputs=x_input, outputs=z)`` (model/model.py:218): x -> MultivariateNormalTriL."""

    def __init__(self, net: Deblender):
        self.net = net

    def __call__(self, x, eps=None, sample=True, seed=None):
        p = self.net.encode(x)
        z, loc, std = self.net.latent(p, eps=eps, sample=sample, seed=seed)
        return MVNTriLOutput(loc, std, z)


def _resolve_weights(survey, weights):
    if isinstance(weights, dict):
        return weights
    if isinstance(weights, str) and weights.startswith("random"):
        seed = int(weights.split(":", 1)[1]) if ":" in weights else 1234
        return spec.random_weights(seed)
    if isinstance(weights, str) and weights.endswith(".npz"):
        with np.load(weights) as f:
            return {k: f[k] for k in f.files}
    dirs = []
    if isinstance(weights, str):
        dirs.append(weights)
    if os.environ.get("DEBVADER_WEIGHTS_DIR"):
        dirs.append(os.path.join(os.environ["DEBVADER_WEIGHTS_DIR"], survey))
    dirs.append(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "data", "weights", survey))
    errors = []
    for d in dirs:
        print(d)  # the reference prints the loading path (model/model.py:264)
        latest = ckpt.latest_checkpoint(d) if os.path.isdir(d) else (d if os.path.exists(d + ".index") else None)
        if latest is None:
            errors.append(f"{d}: no checkpoint")
            continue
        try:
            return ckpt.load_checkpoint(latest)
        except FileNotFoundError as e:  # index present, tensor data missing
            errors.append(str(e))
    raise FileNotFoundError(
        f"no usable '{survey}' checkpoint found ({'; '.join(errors)}). Pass weights=<dict | .npz | checkpoint dir>, "
        "set DEBVADER_WEIGHTS_DIR, or use weights='random[:seed]' for random-init weights of the same architecture."
    )


def create_model_vae(input_shape, latent_dim, filters, kernels, conv_activation=None, dense_activation=None, for_onnx=False,
                     *, weights="random", precision=None, device=None, chunk=0, seed=0):
    """Reference signature model/model.py:164-172; returns (net, encoder, decoder, z-model)."""
    if not spec.is_dc2(input_shape, latent_dim, filters, kernels):
        raise NotImplementedError(
            "debvader_b200 is specialised for the DC2 deblender: input_shape=(59,59,6), latent_dim=32, "
            f"filters=[32,64,128,256], kernels=[3,3,3,3]; got {input_shape}, {latent_dim}, {list(filters)}, {list(kernels)}"
        )
    if conv_activation is not None or dense_activation is not None:
        raise NotImplementedError("only the reference's activation=None configuration is implemented")
    net = Deblender(_resolve_weights("dc2", weights), precision=precision, device=device, chunk=chunk, seed=seed)
    return net, EncoderModel(net), DecoderModel(net), LatentModel(net)


def load_deblender(survey, input_shape, latent_dim, filters, kernels, return_encoder_decoder_z=False, for_onnx=False,
                   *, weights=None, precision=None, device=None, chunk=0, seed=0):
    """Reference signature model/model.py:221-229.  Extension kwargs are keyword-only:
    weights (dict | .npz | checkpoint dir | 'random[:seed]'), precision ('mixed'|'bf16x3'|'fp16x3'|'bf16'|'fp32'), device, chunk, seed."""
    if not spec.is_dc2(input_shape, latent_dim, filters, kernels):
        raise NotImplementedError("debvader_b200 implements the DC2 deblender architecture only (no fallback)")
    net = Deblender(_resolve_weights(survey, weights), precision=precision, device=device, chunk=chunk, seed=seed)
    if return_encoder_decoder_z:
        return net, EncoderModel(net), DecoderModel(net), LatentModel(net)
    return net
