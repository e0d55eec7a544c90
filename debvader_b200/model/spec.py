"""Architecture table of the DC2 deblender (reference model/model.py:61-161).

Tensor keys are the TF2 object-graph checkpoint keys of the shipped checkpoint
(weights/dc2/weights_noisy_v4.386--6.61.ckpt.index) without the
"/.ATTRIBUTES/VARIABLE_VALUE" suffix.  layer_with_weights-0 = encoder model,
layer_with_weights-1 = decoder model.
"""
from __future__ import annotations

import numpy as np

DC2_INPUT_SHAPE = (59, 59, 6)
DC2_LATENT_DIM = 32
DC2_FILTERS = (32, 64, 128, 256)
DC2_KERNELS = (3, 3, 3, 3)

ENC = "layer_with_weights-0/layer_with_weights-%d/%s"
DEC = "layer_with_weights-1/layer_with_weights-%d/%s"


def params_size(latent_dim: int) -> int:
    """tfp.layers.MultivariateNormalTriL.params_size (model/model.py:96-98)."""
    return latent_dim + latent_dim * (latent_dim + 1) // 2


def is_dc2(input_shape, latent_dim, filters, kernels) -> bool:
    return (
        tuple(input_shape) == DC2_INPUT_SHAPE
        and int(latent_dim) == DC2_LATENT_DIM
        and tuple(filters) == DC2_FILTERS
        and tuple(kernels) == DC2_KERNELS
    )


def tensor_table():
    """[(key, shape)] for the 64 model tensors, in checkpoint-index order of layers."""
    H, _, C = DC2_INPUT_SHAPE
    t = [(ENC % (0, n), (C,)) for n in ("gamma", "beta", "moving_mean", "moving_variance")]
    n, h, cin = 1, H, C
    for f in DC2_FILTERS:
        h2 = -(-h // 2)
        t += [
            (ENC % (n, "kernel"), (3, 3, cin, f)),
            (ENC % (n, "bias"), (f,)),
            (ENC % (n + 1, "alpha"), (h, h, f)),
            (ENC % (n + 2, "kernel"), (3, 3, f, f)),
            (ENC % (n + 2, "bias"), (f,)),
            (ENC % (n + 3, "alpha"), (h2, h2, f)),
        ]
        n, h, cin = n + 4, h2, f
    flat = h * h * cin
    P = params_size(DC2_LATENT_DIM)
    t += [(ENC % (n, "alpha"), (flat,)), (ENC % (n + 1, "kernel"), (flat, P)), (ENC % (n + 1, "bias"), (P,))]
    w = int(np.ceil(H / 2 ** len(DC2_FILTERS)))
    t += [
        (DEC % (0, "alpha"), (DC2_LATENT_DIM,)),
        (DEC % (1, "kernel"), (DC2_LATENT_DIM, P)),
        (DEC % (1, "bias"), (P,)),
        (DEC % (2, "alpha"), (P,)),
        (DEC % (3, "kernel"), (P, w * w * DC2_FILTERS[-1])),
        (DEC % (3, "bias"), (w * w * DC2_FILTERS[-1],)),
        (DEC % (4, "alpha"), (w * w * DC2_FILTERS[-1],)),
    ]
    n, h, cin = 5, w, DC2_FILTERS[-1]
    for f in reversed(DC2_FILTERS):
        t += [
            (DEC % (n, "kernel"), (3, 3, f, cin)),
            (DEC % (n, "bias"), (f,)),
            (DEC % (n + 1, "alpha"), (2 * h, 2 * h, f)),
            (DEC % (n + 2, "kernel"), (3, 3, f, f)),
            (DEC % (n + 2, "bias"), (f,)),
            (DEC % (n + 3, "alpha"), (2 * h, 2 * h, f)),
        ]
        n, h, cin = n + 4, 2 * h, f
    t += [(DEC % (n, "kernel"), (3, 3, cin, 2 * C)), (DEC % (n, "bias"), (2 * C,))]
    return t


def _pfx(fmt, n):
    return (fmt % (n, ""))[:-1]  # "layer_with_weights-a/layer_with_weights-n"


def layer_types():
    """[(checkpoint key prefix, Keras layer type)] of every weighted layer in order — the layer sequence the library and the
    oracle implement (reference model/model.py:61-161): BatchNormalization, then per filter Conv2D / PReLU / Conv2D(stride 2)
    / PReLU, the Flatten PReLU and the Dense head; decoder PReLU, Dense, PReLU, Dense, PReLU, per filter
    Conv2DTranspose(stride 2) / PReLU / Conv2DTranspose / PReLU, and the Conv2D output layer.  Pinned by the object graph of
    the shipped checkpoint (tests/golden/object_graph_names.json)."""
    t = [(_pfx(ENC, 0), "batch_normalization")]
    n = 1
    for _ in DC2_FILTERS:
        t += [(_pfx(ENC, n), "conv2d"), (_pfx(ENC, n + 1), "p_re_lu"), (_pfx(ENC, n + 2), "conv2d"), (_pfx(ENC, n + 3), "p_re_lu")]
        n += 4
    t += [(_pfx(ENC, n), "p_re_lu"), (_pfx(ENC, n + 1), "dense")]
    t += [(_pfx(DEC, 0), "p_re_lu"), (_pfx(DEC, 1), "dense"), (_pfx(DEC, 2), "p_re_lu"), (_pfx(DEC, 3), "dense"),
          (_pfx(DEC, 4), "p_re_lu")]
    n = 5
    for _ in DC2_FILTERS:
        t += [(_pfx(DEC, n), "conv2d_transpose"), (_pfx(DEC, n + 1), "p_re_lu"), (_pfx(DEC, n + 2), "conv2d_transpose"),
              (_pfx(DEC, n + 3), "p_re_lu")]
        n += 4
    t += [(_pfx(DEC, n), "conv2d")]
    return t


FLOP_PER_STAMP = 658_693_504  # SURVEY §2.4 / BASELINE.md: nominal 2*MACs, encoder 191 648 128 + decoder 467 045 376


def random_weights(seed: int = 1234, dtype=np.float32):
    """Seeded random-init weights of the DC2 architecture (north_star allows random init:
    the shipped checkpoint's data shard is not distributed with the reference snapshot).
    glorot-uniform x1.6 kernels, N(0,.05) biases, PReLU alpha ~ U(0,.25), BN gamma ~ U(.5,1.5),
    beta ~ N(0,.1), moving_mean ~ N(.05,.02), moving_variance ~ U(.05,.15)."""
    rng = np.random.default_rng(seed)
    w = {}
    for key, shape in tensor_table():
        name = key.rsplit("/", 1)[1]
        if name == "kernel":
            if len(shape) == 4:
                fan_in, fan_out = 9 * shape[2], 9 * shape[3]
            else:
                fan_in, fan_out = shape
            lim = np.sqrt(6.0 / (fan_in + fan_out))
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
        else:  # moving_variance
            v = rng.uniform(0.05, 0.15, size=shape)
        w[key] = np.ascontiguousarray(v.astype(dtype))
    return w


# nominal MACs per stamp of each GEMM-shaped layer (padded taps counted), keyed by the
# layer names the library reports in dbv_layer_times
LAYER_MACS = {
    "enc_conv1": 6_015_168, "enc_conv2": 8_294_400, "enc_conv3": 16_588_800, "enc_conv4": 8_294_400,
    "enc_conv5": 16_588_800, "enc_conv6": 9_437_184, "enc_conv7": 18_874_368, "enc_conv8": 9_437_184,
    "enc_dense": 2_293_760, "latent": 0, "enc_im2col": 0, "dec_dense1": 17_920, "dec_dense2": 2_293_760,
    "dec_convT1": 9_437_184, "dec_convT2": 37_748_736, "dec_convT3": 18_874_368, "dec_convT4": 37_748_736,
    "dec_convT5": 18_874_368, "dec_convT6": 37_748_736, "dec_convT7": 18_874_368, "dec_convT8": 37_748_736,
    "dec_head": 14_155_776,
}
assert 2 * sum(LAYER_MACS.values()) == FLOP_PER_STAMP
