#!/usr/bin/env python
"""Pin the network oracle against the REAL reference.

Needs an environment where the reference's dependencies import (tensorflow==2.13.0, tensorflow-probability==0.21.0,
requirements.txt:9-10) and the reference source is reachable (``--reference /path/to/debvader/src`` or an installed
``debvader``).  Neither is the case in the build container or on the GPU box (no network, no TF wheel), so this script
has NOT been run there: until it has, network parity stays "unpinned" (DESIGN.md section 2).

What it does
  1. builds the Keras models with the reference's own ``create_model_vae`` (model/model.py:164-218),
  2. ``set_weights`` from ``oracle.weights.make_random_weights(seed)`` in checkpoint order (the order of
     ``layer_table()`` is the order of ``model.weights`` of the encoder, then the decoder — checked by shape),
  3. runs ``encoder(x)``; computes ``z = loc + L eps`` from those params with the reference's explicit twin of the TFP layer
     (``MvNormal``, model/model.py:43-58, with eps supplied instead of drawn); runs ``decoder(z)`` and takes
     ``.mean()`` / ``.stddev()`` of the returned distribution,
  4. compares every stage with the numpy oracle (``oracle/vae_numpy.py``) on the same inputs and prints max abs errors
     relative to the peak flux,
  5. writes ``tests/golden/network_tf.npz`` (inputs, eps, seed and TF's outputs).  ``tests/test_oracle_network.py::
     test_oracle_matches_tensorflow_golden`` consumes that file when it exists, and ``tests/test_gpu_network.py`` then
     compares the CUDA path with TensorFlow's numbers directly.

    python tools/tf_crosscheck.py [--reference /root/reference/src] [--seed 1234] [--stamps 8] [--out tests/golden/network_tf.npz]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference/src")
    ap.add_argument("--seed", type=int, default=1234)
    ap.add_argument("--stamps", type=int, default=8)
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden", "network_tf.npz"))
    args = ap.parse_args()

    try:
        import tensorflow as tf
        import tensorflow_probability as tfp  # noqa: F401
    except Exception as e:  # the expected outcome in the build container
        print(f"tensorflow / tensorflow_probability not importable here ({type(e).__name__}: {e}); nothing written. "
              "Run this where the reference's requirements.txt is installed.")
        return 2
    if args.reference and os.path.isdir(args.reference):
        sys.path.insert(0, args.reference)
    import types

    sys.modules.setdefault("sep", types.ModuleType("sep"))  # debvader/__init__ imports the detector; not needed here
    from debvader.model.model import create_model_vae  # the reference's own graph

    from oracle import vae_numpy as vn
    from oracle import weights as ow

    cfg = ((59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3])
    net, encoder, decoder, zmodel = create_model_vae(*cfg)
    w = ow.make_random_weights(seed=args.seed)
    table = ow.layer_table()
    enc_keys = [k for k, _ in table if k.startswith("layer_with_weights-0/")]
    dec_keys = [k for k, _ in table if k.startswith("layer_with_weights-1/")]

    def assign(model, keys):
        # checkpoint keys are "layer_with_weights-<n>/<attr>": the n-th layer of the model that owns variables, attribute by name
        layers = [l for l in model.layers if l.weights]
        by_layer = {}
        for k in keys:
            n = int(k.split("/")[1].rsplit("-", 1)[1])
            by_layer.setdefault(n, {})[k.rsplit("/", 1)[1]] = w[k]
        assert len(layers) == len(by_layer), (len(layers), len(by_layer))
        for n, layer in enumerate(layers):
            vals = []
            for v in layer.weights:
                name = v.name.split("/")[-1].split(":")[0]
                a = by_layer[n][name]
                assert tuple(v.shape) == a.shape, (layer.name, name, tuple(v.shape), a.shape)
                vals.append(a)
            layer.set_weights(vals)

    assign(encoder, enc_keys)
    assign(decoder, dec_keys)

    x = ow.synthetic_stamps(args.stamps, seed=11).astype(np.float32)
    eps = np.random.default_rng(0).normal(size=(args.stamps, 32)).astype(np.float32)
    params_tf = encoder(tf.constant(x)).numpy()  # inference mode, as deblend() calls net(x) (deblender.py:18)
    # the reference's explicit twin of the TFP layer (model/model.py:43-58) with eps supplied
    from tensorflow_probability.python.math import fill_triangular

    t = tf.constant(params_tf)
    scale_tril = fill_triangular(t[..., 32:])
    diag = tf.nn.softplus(tf.linalg.diag_part(scale_tril)) + np.float32(1e-5)
    scale_tril = tf.linalg.set_diag(scale_tril, diag)
    z_tf = (t[..., :32] + tf.linalg.matvec(scale_tril, tf.constant(eps))).numpy()
    # ... and the TFP layer itself must describe the same distribution
    dist_z = zmodel(tf.constant(x))
    assert np.allclose(dist_z.mean().numpy(), params_tf[:, :32], atol=1e-6)
    assert np.allclose(dist_z.stddev().numpy(), np.sqrt((scale_tril.numpy() ** 2).sum(-1)), rtol=1e-5, atol=1e-6)
    out = decoder(tf.constant(z_tf))
    mean_tf, std_tf = out.mean().numpy(), out.stddev().numpy()

    o = vn.forward(w, x.astype(np.float64), eps.astype(np.float64))
    peak = float(np.abs(mean_tf).max())
    report = {}
    for name, a, b in (("params", params_tf, o["params"]), ("z", z_tf, o["z"]), ("mean", mean_tf, o["mean"]), ("stddev", std_tf, o["stddev"])):
        report[name] = float(np.abs(a.astype(np.float64) - np.asarray(b, dtype=np.float64)).max())
    print("max |TF - oracle|:", report, "peak flux", peak, "-> mean err / peak = %.3e" % (report["mean"] / peak))
    np.savez_compressed(args.out, seed=args.seed, x=x, eps=eps, params=params_tf, z=z_tf, mean=mean_tf, stddev=std_tf,
                        tf_version=tf.__version__)
    print("wrote", args.out)
    ok = report["mean"] / peak <= 1e-5 and report["stddev"] / peak <= 1e-5
    print("oracle pinned to TensorFlow" if ok else "MISMATCH: the oracle's [ext] rules need fixing")
    return 0 if ok else 1


if __name__ == "__main__":
    sys.exit(main())
