"""numpy restatement of the reference network (oracle; test infrastructure).

Follows /root/reference/src/debvader/model/model.py:
  * encoder  ``create_encoder``    model.py:61-100
  * latent   ``MvNormal.__call__`` model.py:43-58 (the reference's own explicit
             twin of ``tfp.layers.MultivariateNormalTriL(32)`` at model.py:211-214)
  * decoder  ``create_decoder``    model.py:103-161

Semantics that live in un-vendored third-party code (TensorFlow 2.13.0, Keras,
tensorflow-probability 0.21.0 — requirements.txt:9-10) are restated from their
published behaviour and flagged [ext]:
  [ext] Keras ``BatchNormalization()`` in inference: axis=-1, epsilon=1e-3,
        y = gamma*(x-mean)/sqrt(var+eps)+beta.
  [ext] TF "SAME" padding: out=ceil(in/s); pad_total=max((out-1)*s+k-in,0);
        before=pad_total//2, after=pad_total-before.  Conv2D is a
        cross-correlation, NHWC x HWIO.
  [ext] Keras ``PReLU()`` default shared_axes=None: one alpha per (h,w,c);
        f(x)=max(x,0)+alpha*min(x,0).
  [ext] ``Conv2DTranspose(padding='same')`` = gradient of a SAME conv whose
        *input* has the transposed conv's output size s*n; kernel layout
        (kh,kw,out,in);  out[y,x,co] = b[co] + sum in[i,j,ci]*W[ky,kx,co,ci]
        over s*i+ky-pb==y, s*j+kx-pb==x with pb = max(k-s,0)//2.
  [ext] ``fill_triangular`` (lower): xc=concat(x[n:], reverse(x)) -> (n,n)
        row-major -> lower band.
  [ext] ``tf.nn.softplus`` = log1p(exp(x)).

Everything is plain numpy at the dtype of the inputs (float64 for the
high-precision oracle, float32 to imitate TF's default arithmetic).  It is
O(taps) einsums per layer: fine for tens of stamps, not a benchmark.
"""
from __future__ import annotations

import numpy as np

BN_EPS = 1e-3  # [ext] Keras BatchNormalization default epsilon
DIAG_SHIFT = 1e-5  # model.py:49
SCALE_SHIFT = 1e-4  # model.py:155-157

E = "layer_with_weights-0/layer_with_weights-%d/%s"
D = "layer_with_weights-1/layer_with_weights-%d/%s"


def same_pad(n: int, k: int, s: int):
    out = -(-n // s)
    total = max((out - 1) * s + k - n, 0)
    before = total // 2
    return out, before, total - before


def batchnorm_inference(x, gamma, beta, mean, var):
    """model.py:79 [ext]."""
    return gamma * (x - mean) / np.sqrt(var + x.dtype.type(BN_EPS)) + beta


def conv2d_same(x, w, b, stride):
    """Conv2D(padding='same', strides=stride) — model.py:80-91, 137.  x NHWC, w (kh,kw,ci,co)."""
    B, H, W, _ = x.shape
    kh, kw, _, co = w.shape
    Ho, pt, pb = same_pad(H, kh, stride)
    Wo, pl, pr = same_pad(W, kw, stride)
    xp = np.pad(x, ((0, 0), (pt, pb), (pl, pr), (0, 0)))
    y = np.zeros((B, Ho, Wo, co), dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            win = xp[:, ky : ky + stride * (Ho - 1) + 1 : stride, kx : kx + stride * (Wo - 1) + 1 : stride, :]
            y += win @ w[ky, kx]
    return y + b


def conv2d_transpose_same(x, w, b, stride):
    """Conv2DTranspose(padding='same', strides=stride) — model.py:120-135 [ext].  w (kh,kw,co,ci)."""
    B, H, W, _ = x.shape
    kh, kw, co, _ = w.shape
    pbh = max(kh - stride, 0) // 2
    pbw = max(kw - stride, 0) // 2
    full = np.zeros((B, (H - 1) * stride + kh, (W - 1) * stride + kw, co), dtype=x.dtype)
    for ky in range(kh):
        for kx in range(kw):
            full[:, ky : ky + stride * (H - 1) + 1 : stride, kx : kx + stride * (W - 1) + 1 : stride, :] += x @ w[ky, kx].T
    Ho, Wo = stride * H, stride * W
    out = np.zeros((B, Ho, Wo, co), dtype=x.dtype)
    # rows [pb, pb+Ho) of `full`, clipped to what exists (full has (H-1)s+k rows)
    hh = min(Ho, full.shape[1] - pbh)
    ww = min(Wo, full.shape[2] - pbw)
    out[:, :hh, :ww] = full[:, pbh : pbh + hh, pbw : pbw + ww]
    return out + b


def prelu(x, alpha):
    """Keras PReLU [ext]: max(x,0) + alpha*min(x,0)."""
    return np.maximum(x, 0) + alpha * np.minimum(x, 0)


def softplus(x):
    return np.logaddexp(x, x.dtype.type(0))


def fill_triangular_lower(x):
    """tfp.math.fill_triangular(x, upper=False) [ext]; x (..., n(n+1)/2) -> (..., n, n)."""
    m = x.shape[-1]
    n = int((np.sqrt(8 * m + 1) - 1) / 2)
    assert n * (n + 1) // 2 == m
    xc = np.concatenate([x[..., n:], x[..., ::-1]], axis=-1)
    y = xc.reshape(x.shape[:-1] + (n, n))
    return np.tril(y)


def encode(wts, x):
    """create_encoder, model.py:61-100.  x (B,59,59,6) -> params (B,560)."""
    dt = x.dtype
    g = lambda k: wts[k].astype(dt)
    h = batchnorm_inference(x, g(E % (0, "gamma")), g(E % (0, "beta")), g(E % (0, "moving_mean")), g(E % (0, "moving_variance")))
    n = 1
    while (E % (n, "kernel")) in wts and wts[E % (n, "kernel")].ndim == 4:
        h = conv2d_same(h, g(E % (n, "kernel")), g(E % (n, "bias")), 1)
        h = prelu(h, g(E % (n + 1, "alpha")))
        h = conv2d_same(h, g(E % (n + 2, "kernel")), g(E % (n + 2, "bias")), 2)
        h = prelu(h, g(E % (n + 3, "alpha")))
        n += 4
    h = h.reshape(h.shape[0], -1)  # Flatten, (h,w,c) order — model.py:94
    h = prelu(h, g(E % (n, "alpha")))  # model.py:95
    return h @ g(E % (n + 1, "kernel")) + g(E % (n + 1, "bias"))  # model.py:96-98


def latent(params, eps=None, latent_dim=32):
    """MvNormal.__call__, model.py:48-58.  Returns dict(loc, scale_tril, z, stddev).

    ``eps`` (B,latent) is the N(0,1) draw of model.py:57; None means eps=0 (z=loc).
    ``stddev`` is what tfd.MultivariateNormalTriL.stddev() returns [ext]:
    sqrt(sum_j L[i,j]^2).
    """
    dt = params.dtype
    loc = params[..., :latent_dim]
    tril = fill_triangular_lower(params[..., latent_dim:])
    idx = np.arange(latent_dim)
    tril[..., idx, idx] = softplus(tril[..., idx, idx]) + dt.type(DIAG_SHIFT)
    if eps is None:
        z = loc.copy()
    else:
        z = loc + np.einsum("...ij,...j->...i", tril, eps.astype(dt))
    return {"loc": loc, "scale_tril": tril, "z": z, "stddev": np.sqrt(np.sum(tril * tril, axis=-1))}


def decode(wts, z, input_shape=(59, 59, 6)):
    """create_decoder, model.py:103-161.  z (B,32) -> (mean, stddev) each (B,59,59,6)."""
    dt = z.dtype
    g = lambda k: wts[k].astype(dt)
    h = prelu(z, g(D % (0, "alpha")))  # model.py:113
    h = prelu(h @ g(D % (1, "kernel")) + g(D % (1, "bias")), g(D % (2, "alpha")))  # :114-115
    h = prelu(h @ g(D % (3, "kernel")) + g(D % (3, "bias")), g(D % (4, "alpha")))  # :117-118
    n = 5
    cin = wts[D % (n, "kernel")].shape[3]
    w = int(round(np.sqrt(h.shape[1] // cin)))
    h = h.reshape(-1, w, w, cin)  # model.py:119
    while wts[D % (n, "kernel")].shape[3] == h.shape[-1] and (D % (n + 1, "alpha")) in wts:
        h = conv2d_transpose_same(h, g(D % (n, "kernel")), g(D % (n, "bias")), 2)
        h = prelu(h, g(D % (n + 1, "alpha")))
        h = conv2d_transpose_same(h, g(D % (n + 2, "kernel")), g(D % (n + 2, "bias")), 1)
        h = prelu(h, g(D % (n + 3, "alpha")))
        n += 4
    h = np.maximum(conv2d_same(h, g(D % (n, "kernel")), g(D % (n, "bias")), 1), 0)  # :137 relu
    crop = h.shape[1] - input_shape[0]  # model.py:140-148
    if crop > 0:
        lo = crop // 2
        hi = crop - lo if crop % 2 else lo
        h = h[:, lo : h.shape[1] - hi, lo : h.shape[2] - hi, :]
    C = input_shape[-1]
    return h[..., :C], dt.type(SCALE_SHIFT) + h[..., C:]  # model.py:155-157


def forward(wts, x, eps=None):
    """net(x) of create_model_vae (model.py:216) with the latent draw made explicit."""
    params = encode(wts, x)
    lat = latent(params, eps)
    mean, std = decode(wts, lat["z"])
    return {"params": params, "z": lat["z"], "z_loc": lat["loc"], "z_stddev": lat["stddev"], "mean": mean, "stddev": std}
