"""Self-consistency of the network oracle: explicit numpy restatement vs torch-CPU library convs."""
import numpy as np
import pytest
import torch

from oracle import vae_numpy as vn
from oracle import weights as ow
from oracle.vae_torch import TorchOracle


@pytest.fixture(scope="module")
def wts():
    return ow.make_random_weights(seed=1234)


def test_same_pad_rule():
    # 59->30 pad (1,1); 30->15 pad (0,1); 15->8 pad (1,1); 8->4 pad (0,1)
    assert vn.same_pad(59, 3, 2) == (30, 1, 1)
    assert vn.same_pad(30, 3, 2) == (15, 0, 1)
    assert vn.same_pad(15, 3, 2) == (8, 1, 1)
    assert vn.same_pad(8, 3, 2) == (4, 0, 1)
    assert vn.same_pad(59, 3, 1) == (59, 1, 1)


def test_fill_triangular_index_form():
    # rows 0-15 L[i,j]=t[64+32i+j]; rows 16-31 L[i,j]=t[1055-32i-j]
    t = np.arange(560, dtype=np.float64)
    L = vn.fill_triangular_lower(t[32:])
    for i in range(32):
        for j in range(i + 1):
            exp = t[64 + 32 * i + j] if i < 16 else t[1055 - 32 * i - j]
            assert L[i, j] == exp
    assert np.all(np.triu(L, 1) == 0)


def test_fill_triangular_is_the_op_chain_tensorflow_recorded():
    """The reference's own notebook (notebooks/deblender_to_onnx.ipynb, output of `net.summary()` for the for_onnx model,
    model/model.py:43-58) lists the TensorFlow ops fill_triangular expanded into, with their shapes:
    getitem (None, 528) = t[..., 32:]  ->  getitem_3 (None, 496) = x[..., n:]  +  reverse (None, 528)  ->  concat_1 (None, 1024)
    ->  reshape (None, 32, 32)  ->  linalg.band_part (lower)  ->  diag_part / softplus / add / set_diag (None, 32, 32)
    ->  matmul with the (None, 32, 1) draw  ->  add to t[..., :32].  The oracle's restatement is that chain, step by step."""
    rng = np.random.default_rng(0)
    t = rng.normal(size=(3, 560))
    x = t[..., 32:]
    assert x.shape == (3, 528)
    n = 32
    head, rev = x[..., n:], x[..., ::-1]
    assert head.shape == (3, 496) and rev.shape == (3, 528)
    cat = np.concatenate([head, rev], axis=-1)
    assert cat.shape == (3, 1024)
    lower = np.tril(cat.reshape(3, 32, 32))  # band_part(num_lower=-1, num_upper=0)
    assert np.array_equal(vn.fill_triangular_lower(x), lower)
    diag = np.log1p(np.exp(np.diagonal(lower, axis1=-2, axis2=-1))) + 1e-5  # softplus + diag_shift (model.py:50-52)
    L = lower.copy()
    L[:, np.arange(32), np.arange(32)] = diag
    eps = rng.normal(size=(3, 32))
    want = t[..., :32] + np.einsum("bij,bj->bi", L, eps)  # loc + matvec(scale_tril, samples)
    got = vn.latent(t, eps)
    got_z = got["z"] if isinstance(got, dict) else got[0]
    np.testing.assert_allclose(got_z, want, rtol=1e-12, atol=1e-12)


def test_transposed_conv_tiny_bruteforce():
    rng = np.random.default_rng(0)
    for s in (1, 2):
        x = rng.normal(size=(1, 3, 3, 2))
        w = rng.normal(size=(3, 3, 4, 2))
        b = rng.normal(size=4)
        y = vn.conv2d_transpose_same(x, w, b, s)
        pb = 0 if s == 2 else 1
        ref = np.zeros((1, 3 * s, 3 * s, 4)) + b
        for i in range(3):
            for j in range(3):
                for ky in range(3):
                    for kx in range(3):
                        yy, xx = s * i + ky - pb, s * j + kx - pb
                        if 0 <= yy < 3 * s and 0 <= xx < 3 * s:
                            ref[0, yy, xx] += w[ky, kx] @ x[0, i, j]
        np.testing.assert_allclose(y, ref, atol=1e-12)


def test_numpy_vs_torch_fp64(wts):
    x = ow.synthetic_stamps(3, seed=5, dtype=np.float64)
    eps = np.random.default_rng(1).normal(size=(3, 32))
    a = vn.forward(wts, x, eps)
    b = TorchOracle(wts, dtype=torch.float64).forward(x, eps)
    for k in ("params", "z", "z_stddev", "mean", "stddev"):
        np.testing.assert_allclose(a[k], b[k].numpy(), rtol=1e-9, atol=1e-9, err_msg=k)
    assert a["mean"].shape == (3, 59, 59, 6) and a["stddev"].shape == (3, 59, 59, 6)
    assert a["mean"].min() >= 0 and a["stddev"].min() >= 1e-4
    assert a["mean"].max() > 0.05, "random init must produce a non-trivial output"


def test_fp32_vs_fp64_within_north_star_tolerance(wts):
    x = ow.synthetic_stamps(4, seed=6)
    eps = np.random.default_rng(2).normal(size=(4, 32)).astype(np.float32)
    a = TorchOracle(wts, dtype=torch.float64).forward(x.astype(np.float64), eps.astype(np.float64))
    b = TorchOracle(wts, dtype=torch.float32).forward(x, eps)
    peak = float(a["mean"].abs().max())
    err = float((a["mean"] - b["mean"].double()).abs().max())
    assert err <= 1e-5 * max(peak, 1.0), (err, peak)


def test_oracle_matches_tensorflow_golden(golden_dir):
    """Consumes tests/golden/network_tf.npz — outputs of the REAL reference (Keras/TFP) written by tools/tf_crosscheck.py
    wherever TensorFlow 2.13 is installed.  The file cannot be produced in the build container (no TensorFlow), so until
    somebody runs that script the network oracle stays "parity unpinned" and this test is skipped."""
    import os

    path = os.path.join(golden_dir, "network_tf.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/network_tf.npz absent: tools/tf_crosscheck.py has not been run where TensorFlow exists (network parity unpinned)")
    g = np.load(path)
    w = ow.make_random_weights(seed=int(g["seed"]))
    o = vn.forward(w, g["x"].astype(np.float64), g["eps"].astype(np.float64))
    peak = float(np.abs(g["mean"]).max())
    for k in ("params", "z"):
        np.testing.assert_allclose(o[k], g[k], rtol=0, atol=2e-5 * max(1.0, float(np.abs(g[k]).max())), err_msg=k)
    assert float(np.abs(o["mean"] - g["mean"]).max()) <= 1e-5 * peak  # TF computes in fp32: the north_star's fp32 tolerance
    assert float(np.abs(o["stddev"] - g["stddev"]).max()) <= 1e-5 * peak
