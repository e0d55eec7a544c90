"""Weight naming + seeded random-init generator (oracle side; test infrastructure).

Keys are the TF2 object-graph checkpoint keys of the shipped DC2 checkpoint
(``src/debvader/data/weights/dc2/weights_noisy_v4.386--6.61.ckpt.index``) with
the ``/.ATTRIBUTES/VARIABLE_VALUE`` suffix stripped, e.g.
``layer_with_weights-0/layer_with_weights-1/kernel``.  ``layer_with_weights-0``
is the encoder model (reference model/model.py:61-100), ``layer_with_weights-1``
the decoder model (model/model.py:103-161).

The product package carries its own copy of the layer table
(``debvader_b200/model/spec.py``); tests assert both agree.
"""
from __future__ import annotations

import numpy as np

INPUT_SHAPE = (59, 59, 6)
LATENT_DIM = 32
FILTERS = (32, 64, 128, 256)
KERNELS = (3, 3, 3, 3)


def params_size(latent_dim: int) -> int:
    """tfp.layers.MultivariateNormalTriL.params_size (model/model.py:96-98)."""
    return latent_dim + latent_dim * (latent_dim + 1) // 2


def same_out(n: int, stride: int) -> int:
    return -(-n // stride)


def layer_table(input_shape=INPUT_SHAPE, latent_dim=LATENT_DIM, filters=FILTERS, kernels=KERNELS):
    """Ordered list of (key, shape) for all 64 model tensors.

    Encoder: model/model.py:79-98.  Decoder: model/model.py:113-137.
    """
    H, W, C = input_shape
    assert H == W
    out = []
    e = "layer_with_weights-0/layer_with_weights-%d/%s"
    d = "layer_with_weights-1/layer_with_weights-%d/%s"
    for nm in ("gamma", "beta", "moving_mean", "moving_variance"):
        out.append((e % (0, nm), (C,)))
    n = 1
    h, cin = H, C
    for f, k in zip(filters, kernels):
        out.append((e % (n, "kernel"), (k, k, cin, f)))
        out.append((e % (n, "bias"), (f,)))
        out.append((e % (n + 1, "alpha"), (h, h, f)))
        h2 = same_out(h, 2)
        out.append((e % (n + 2, "kernel"), (k, k, f, f)))
        out.append((e % (n + 2, "bias"), (f,)))
        out.append((e % (n + 3, "alpha"), (h2, h2, f)))
        n += 4
        h, cin = h2, f
    flat = h * h * cin
    out.append((e % (n, "alpha"), (flat,)))
    out.append((e % (n + 1, "kernel"), (flat, params_size(latent_dim))))
    out.append((e % (n + 1, "bias"), (params_size(latent_dim),)))

    w = int(np.ceil(H / 2 ** len(filters)))
    p32 = params_size(32)  # the reference hard-codes 32 here (model/model.py:114)
    out.append((d % (0, "alpha"), (latent_dim,)))
    out.append((d % (1, "kernel"), (latent_dim, p32)))
    out.append((d % (1, "bias"), (p32,)))
    out.append((d % (2, "alpha"), (p32,)))
    out.append((d % (3, "kernel"), (p32, w * w * filters[-1])))
    out.append((d % (3, "bias"), (w * w * filters[-1],)))
    out.append((d % (4, "alpha"), (w * w * filters[-1],)))
    n = 5
    h, cin = w, filters[-1]
    for i in range(len(filters) - 1, -1, -1):
        f, k = filters[i], kernels[i]
        out.append((d % (n, "kernel"), (k, k, f, cin)))  # Conv2DTranspose: (kh,kw,out,in)
        out.append((d % (n, "bias"), (f,)))
        out.append((d % (n + 1, "alpha"), (2 * h, 2 * h, f)))
        out.append((d % (n + 2, "kernel"), (k, k, f, f)))
        out.append((d % (n + 2, "bias"), (f,)))
        out.append((d % (n + 3, "alpha"), (2 * h, 2 * h, f)))
        n += 4
        h, cin = 2 * h, f
    out.append((d % (n, "kernel"), (3, 3, cin, 2 * C)))
    out.append((d % (n, "bias"), (2 * C,)))
    return out


def make_random_weights(seed: int = 1234, dtype=np.float32, **cfg):
    """Random-init weights of the reference architecture.

    glorot-uniform kernels, small non-zero biases, PReLU alpha ~ U(0, 0.25)
    (Keras initialises alpha to 0, which would never exercise the alpha path),
    BN gamma ~ U(.5,1.5), beta ~ N(0,.1), moving_mean ~ N(.05,.02),
    moving_variance ~ U(.05,.15).  numpy's PCG64 stream is stable across
    versions, so a seed pins the tensors exactly.
    """
    rng = np.random.default_rng(seed)
    w = {}
    for key, shape in layer_table(**cfg):
        name = key.rsplit("/", 1)[1]
        if name == "kernel":
            if len(shape) == 4:
                kh, kw, a, b = shape
                fan_in, fan_out = kh * kw * a, kh * kw * b
            else:
                fan_in, fan_out = shape
            lim = np.sqrt(6.0 / (fan_in + fan_out))
            # a slightly hot init (x1.6) keeps activations O(1) through 20 layers
            v = rng.uniform(-lim, lim, size=shape) * 1.6
        elif name == "bias":
            v = rng.normal(0.0, 0.05, size=shape)
        elif name == "alpha":
            v = rng.uniform(0.0, 0.25, size=shape)
        elif name == "gamma":
            v = rng.uniform(0.5, 1.5, size=shape)
        elif name == "beta":
            v = rng.normal(0.0, 0.1, size=shape)
        elif name == "moving_mean":
            v = rng.normal(0.05, 0.02, size=shape)
        elif name == "moving_variance":
            v = rng.uniform(0.05, 0.15, size=shape)
        else:  # pragma: no cover
            raise KeyError(key)
        w[key] = np.ascontiguousarray(v.astype(dtype))
    return w


def count_params(weights) -> dict:
    enc = sum(v.size for k, v in weights.items() if k.startswith("layer_with_weights-0/"))
    dec = sum(v.size for k, v in weights.items() if k.startswith("layer_with_weights-1/"))
    return {"encoder": enc, "decoder": dec, "total": enc + dec}


def synthetic_stamps(n: int, seed: int = 0, dtype=np.float32):
    """Synthetic 59x59x6 stamps (SURVEY §8d cfg 2): sky noise N(0,0.3) + 1-3
    elliptical Gaussian blobs, one centred, mimicking imgs_dc2.npy statistics."""
    rng = np.random.default_rng(seed)
    S, C = 59, 6
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float64)
    out = rng.normal(0.0, 0.3, size=(n, S, S, C))
    for i in range(n):
        nb = rng.integers(1, 4)
        for b in range(nb):
            cy, cx = (29.0, 29.0) if b == 0 else rng.uniform(8, 51, size=2)
            peak = 10 ** rng.uniform(-0.5, 1.5)
            sx, sy = rng.uniform(1.5, 5.0, size=2)
            th = rng.uniform(0, np.pi)
            dx, dy = xx - cx, yy - cy
            u = dx * np.cos(th) + dy * np.sin(th)
            v = -dx * np.sin(th) + dy * np.cos(th)
            prof = peak * np.exp(-0.5 * ((u / sx) ** 2 + (v / sy) ** 2))
            sed = rng.uniform(0.3, 1.0, size=C)
            out[i] += prof[:, :, None] * sed[None, None, :]
    return out.astype(dtype)
