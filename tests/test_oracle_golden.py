"""The oracle against the reference's own outputs (tests/golden, made by make_golden.py)."""
import hashlib
import json
import os

import numpy as np
import pytest

from oracle import field_numpy as fo
from oracle import weights as ow


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def test_extraction_matches_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "extraction_cases.json")))
    assert g["shipped_fixture"]["galaxies_from_field_equals_reference_extract"]
    for c in g["cases"]:
        field = np.random.default_rng(c["seed"]).random((1, c["F"], c["F"], c["C"]))
        cut, idx = fo.extract_cutouts(field, c["F"], c["centres"], c["S"], c["C"])
        assert idx == c["list_idx"], c["seed"]
        assert sha(cut) == c["sha256"], c["seed"]


def test_reference_unit_test_cases():
    # same four situations as the reference's tests/test_extraction.py:6-62
    rng = np.random.default_rng(0)
    img = rng.random((1, 15, 15, 3))
    cut, idx = fo.extract_cutouts(img, 15, [[-4, -3]], 5, 3)
    np.testing.assert_array_equal(cut, img[:, 1:6, 2:7])
    cut, idx = fo.extract_cutouts(img, 15, [[5, 5]], 5, 3)
    np.testing.assert_array_equal(cut, img[:, 10:, 10:])
    cut, idx = fo.extract_cutouts(img, 15, [[-5, -5]], 5, 3)
    np.testing.assert_array_equal(cut, img[:, :5, :5])
    cut, idx = fo.extract_cutouts(img, 15, [[6, 6]], 5, 3)
    assert idx == []


@pytest.mark.parametrize("name", ["odd", "even"])
def test_residual_and_predicted_fields_match_reference(golden_dir, name):
    g = np.load(os.path.join(golden_dir, "field_ops.npz"))
    field, pos = g[f"{name}_field"], g[f"{name}_pos"]
    means, stds = g[f"{name}_means"], g[f"{name}_stds"]
    S = means.shape[1]
    res = fo.residual_field(field, means, pos[:, 0], pos[:, 1], cutout_size=S)
    # the reference goes through a cubic-spline ndimage.shift: equal to ~1e-13, not bit-equal
    np.testing.assert_allclose(res, g[f"{name}_residual"], rtol=0, atol=1e-12)
    pf = fo.predicted_fields(field.shape[1], field.shape[3], means, stds, None, pos[:, 0], pos[:, 1], cutout_size=S)
    np.testing.assert_allclose(pf["predicted_mean_field"], g[f"{name}_pred_mean"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(pf["predicted_stddev_field"], g[f"{name}_pred_std"], rtol=0, atol=1e-12)
    assert abs(fo.mse(field, res) - float(g[f"{name}_mse"])) < 1e-12


def test_deblend_field_records_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, "deblend_field_fake.npz"))
    field, centres = g["field"], g["centres"]
    cut, idx = fo.extract_cutouts(field, field.shape[1], centres, 59, 6)
    assert idx == list(g["list_idx"])
    np.testing.assert_array_equal(cut[idx], g["cutouts"])
    m = fo.center_mse(cut[idx], g["mean"])
    np.testing.assert_array_equal(~(m > 2.0), g["passed_cuts"])
    res = fo.residual_field(field, g["mean"], g["dx"], g["dy"])
    np.testing.assert_allclose(res, g["residual"], rtol=0, atol=1e-11)


def test_architecture_matches_checkpoint_index(golden_dir):
    a = json.load(open(os.path.join(golden_dir, "architecture.json")))
    table = ow.layer_table()
    assert len(table) == 64 == len(a["tensors"])
    for key, shape in table:
        assert a["tensors"][key] == list(shape), key
    w = ow.make_random_weights(seed=1)
    assert ow.count_params(w) == a["net_summary_params"]


# ---- sub-pixel placement (scipy.ndimage.shift restatement) ------------------------------------
from oracle import spline_numpy as sp  # noqa: E402


@pytest.mark.parametrize("n,shift", [(40, (0.3, -1.7)), (131, (20.25, -33.5)), (64, (3.0, 0.0)), (64, (-2.999999, 2.5)), (20, (7.6, -8.2)), (50, (0.0, 0.5))])
def test_spline_shift_restatement_equals_scipy(n, shift):
    ndi = pytest.importorskip("scipy.ndimage")
    a = np.random.default_rng(n).standard_normal((n, n))
    np.testing.assert_allclose(sp.shift_cubic_constant(a, shift), ndi.shift(a, shift), rtol=0, atol=2e-14)


@pytest.mark.parametrize("name", ["win_odd", "win_even", "whole"])
def test_subpixel_fields_match_reference(golden_dir, name):
    """oracle (full-canvas shift per galaxy and band) and the kernel's windowed formulation (tests/cpu_ops.py)
    against the reference's own get_residual_field / get_predicted_field with fractional positions."""
    from tests import cpu_ops

    g = np.load(os.path.join(golden_dir, "subpixel.npz"))
    field, means, stds = g[f"{name}_field"], g[f"{name}_means"], g[f"{name}_stds"]
    pos = g[f"{name}_pos"] + g[f"{name}_shifts"]
    S = means.shape[1]
    res = sp.residual_field_subpixel(field, means, pos[:, 0], pos[:, 1], S)
    np.testing.assert_allclose(res, g[f"{name}_residual"], rtol=0, atol=1e-12)
    pm = sp.predicted_field_subpixel(field.shape[1], field.shape[3], stds, pos[:, 0], pos[:, 1], S)
    np.testing.assert_allclose(pm, g[f"{name}_pred_std"], rtol=0, atol=1e-12)
    win = cpu_ops.windowed_axpy(field, means, pos[:, 0], pos[:, 1], -1.0)
    np.testing.assert_allclose(win, g[f"{name}_residual"], rtol=0, atol=1e-12)
    win = cpu_ops.windowed_axpy(None, means, pos[:, 0], pos[:, 1], 1.0, field_shape=(field.shape[1], field.shape[1], field.shape[3]))
    np.testing.assert_allclose(win, g[f"{name}_pred_mean"], rtol=0, atol=1e-12)


@pytest.mark.parametrize("S", [59, 64, 7, 2])
def test_warp_scan_prefilter_equals_the_sequential_recursion(S):
    """the scan formulation of spline_place_warp_kernel (tests/cpu_ops.py:warp_scan_prefilter) against scipy's own prefilter
    of a long zero canvas holding the line: coefficients inside the data, and the closed-form head / tail outside."""
    from tests import cpu_ops

    ndi = pytest.importorskip("scipy.ndimage")
    rng = np.random.default_rng(S)
    x = rng.standard_normal(S) * 10
    pad = 60
    canvas = np.zeros(S + 2 * pad)
    canvas[pad : pad + S] = x
    want = ndi.spline_filter1d(canvas, order=3, mode="mirror")
    c, cfirst, cplast = cpu_ops.warp_scan_prefilter(6.0 * x)
    np.testing.assert_allclose(c, want[pad : pad + S], rtol=0, atol=1e-13 * np.abs(want).max())
    z = sp.POLE
    kappa = z / (z * z - 1.0)
    for m in (1, 2, 5, 28):
        assert abs(z**m * cfirst - want[pad - m]) < 1e-13 * np.abs(want).max()
        assert abs(kappa * z**m * cplast - want[pad + S - 1 + m]) < 1e-13 * np.abs(want).max()
