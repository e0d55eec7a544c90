"""-m gpu: the CUDA field operators (through the C-ABI) against the oracle and the reference goldens."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import field_numpy as fo

pytestmark = pytest.mark.gpu


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def ops():
    from debvader_b200 import _fieldops

    return _fieldops


def test_reference_unit_test_cases_through_the_drop_in_path():
    # the four situations of the reference's tests/test_extraction.py:6-62, reference import path
    from debvader.extract.extraction import extract_cutouts

    image = np.random.rand(1, 15, 15, 3)
    cut = extract_cutouts(field_image=image.copy(), field_size=15, galaxy_distances_to_center=[[-4, -3]], cutout_size=5, nb_of_bands=3)
    np.testing.assert_array_equal(cut[0], image[:, 1:6, 2:7])
    cut = extract_cutouts(field_image=image.copy(), field_size=15, galaxy_distances_to_center=[[5, 5]], cutout_size=5, nb_of_bands=3)
    np.testing.assert_array_equal(cut[0], image[:, 10:, 10:])
    cut = extract_cutouts(field_image=image.copy(), field_size=15, galaxy_distances_to_center=[[-5, -5]], cutout_size=5, nb_of_bands=3)
    np.testing.assert_array_equal(cut[0], image[:, :5, :5])
    cut = extract_cutouts(field_image=image.copy(), field_size=15, galaxy_distances_to_center=[[6, 6]], cutout_size=5, nb_of_bands=3)
    assert len(cut[1]) == 0


def test_extract_matches_reference_goldens(golden_dir):
    from debvader_b200.extract.extraction import extract_cutouts

    g = json.load(open(os.path.join(golden_dir, "extraction_cases.json")))
    for c in g["cases"]:
        field = np.random.default_rng(c["seed"]).random((1, c["F"], c["F"], c["C"]))
        cut, idx = extract_cutouts(field, c["F"], c["centres"], c["S"], c["C"])
        assert idx == c["list_idx"], c["seed"]
        assert cut.dtype == np.float64 and sha(cut) == c["sha256"], c["seed"]


def test_extract_dc2_field_golden(golden_dir):
    from debvader_b200.extract.extraction import extract_cutouts

    g = np.load(os.path.join(golden_dir, "dc2_field2.npz"))
    cut, idx = extract_cutouts(g["field"], 259, g["centres"], 59, 6)
    assert idx == list(g["list_idx"])
    assert sha(cut) == str(g["cutouts_sha256"])


@pytest.mark.parametrize("F,S,C,N", [(1025, 59, 6, 700), (300, 59, 5, 64), (128, 7, 1, 100), (64, 9, 3, 50)])
def test_extract_random_vs_oracle(ops, F, S, C, N):
    rng = np.random.default_rng(F + S)
    field = rng.normal(size=(1, F, F, C))
    centres = rng.uniform(-F * 0.7, F * 0.7, size=(N, 2))
    want, widx = fo.extract_cutouts(field, F, centres, S, C)
    plan = ops.plan_windows(centres, S, F)
    fdev = torch.from_numpy(field).cuda()
    got, idx = ops.extract(fdev, plan, S, C, out_dtype=torch.float64)
    assert idx == widx
    np.testing.assert_array_equal(got.cpu().numpy(), want)
    got32, _ = ops.extract(fdev, plan, S, C, out_dtype=torch.float32)  # fused tf.cast(float32)
    np.testing.assert_array_equal(got32.cpu().numpy(), want.astype(np.float32))
    g32, _ = ops.extract(fdev.float(), plan, S, C, out_dtype=torch.float32)
    np.testing.assert_array_equal(g32.cpu().numpy(), want.astype(np.float32))


@pytest.mark.parametrize("F", [259, 260, 77])
def test_window_axpy_bit_exact(ops, F):
    S, C, N = 59, 6, 300
    rng = np.random.default_rng(F)
    field = rng.normal(size=(1, F, F, C))
    pos = rng.integers(-F // 2 - 20, F // 2 + 20, size=(N, 2))  # heavy overlap, some partly / fully outside
    means = rng.random((N, S, S, C)).astype(np.float32)
    stds = rng.random((N, S, S, C)).astype(np.float32)
    want = fo.residual_field(field, means, pos[:, 0], pos[:, 1], cutout_size=S)
    off = ops.subtract_offset(F, S)
    got = ops.window_axpy(torch.from_numpy(field).cuda(), torch.from_numpy(means).cuda(), off + pos[:, 0], off + pos[:, 1], -1.0)
    np.testing.assert_array_equal(got.cpu().numpy(), want)  # bit-exact: same order, one rounding per add
    pf = fo.predicted_fields(F, C, means, stds, None, pos[:, 0], pos[:, 1], cutout_size=S)
    gm = ops.window_axpy(None, torch.from_numpy(stds).cuda(), off + pos[:, 0], off + pos[:, 1], 1.0, field_shape=(F, F, C))
    np.testing.assert_array_equal(gm.cpu().numpy(), pf["predicted_stddev_field"])


def test_window_axpy_matches_reference_spline_path(ops, golden_dir):
    g = np.load(os.path.join(golden_dir, "field_ops.npz"))
    for name in ("odd", "even"):
        field, pos, means = g[f"{name}_field"], g[f"{name}_pos"], g[f"{name}_means"]
        S, F = means.shape[1], field.shape[1]
        off = ops.subtract_offset(F, S)
        got = ops.window_axpy(torch.from_numpy(field).cuda(), torch.from_numpy(means).cuda(), off + pos[:, 0], off + pos[:, 1], -1.0)
        np.testing.assert_allclose(got.cpu().numpy(), g[f"{name}_residual"], rtol=0, atol=1e-12)


def test_window_axpy_more_stamps_than_the_tile_list_holds(ops):
    F, S, C, N = 96, 59, 2, 2000  # every stamp overlaps the central tiles: forces the multi-pass path
    rng = np.random.default_rng(1)
    field = rng.normal(size=(1, F, F, C))
    pos = rng.integers(-6, 7, size=(N, 2))
    means = rng.random((N, S, S, C)).astype(np.float32)
    want = fo.residual_field(field, means, pos[:, 0], pos[:, 1], cutout_size=S)
    off = ops.subtract_offset(F, S)
    got = ops.window_axpy(torch.from_numpy(field).cuda(), torch.from_numpy(means).cuda(), off + pos[:, 0], off + pos[:, 1], -1.0)
    np.testing.assert_array_equal(got.cpu().numpy(), want)


def test_center_mse_and_field_mse(ops):
    rng = np.random.default_rng(2)
    cut = rng.normal(size=(37, 59, 59, 6))
    mean = rng.normal(size=(37, 59, 59, 6)).astype(np.float32)
    want = fo.center_mse(cut, mean)
    got = ops.center_mse(torch.from_numpy(cut).cuda(), torch.from_numpy(mean).cuda(), 24, 34).cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=1e-13)
    a, b = rng.normal(size=(1, 301, 301, 6)), rng.normal(size=(1, 301, 301, 6))
    assert abs(ops.mse(a, b) - fo.mse(a, b)) <= 1e-13 * fo.mse(a, b)
    assert ops.mse(a, a) == 0.0


def test_full_size_round_trip_property(ops):
    """BASELINE cfg 4 size (4096^2 x 6, 2000+ sources): extract -> subtract the very same stamps ->
    every window is exactly zero and everything else is untouched (size-independent property)."""
    F, S, C = 4096, 59, 6
    gen = torch.Generator(device="cuda").manual_seed(0)
    field = torch.randn((1, F, F, C), device="cuda", dtype=torch.float32, generator=gen).double()  # f32-representable values
    g = np.arange(-33, 34)  # 67 x 67 grid of non-overlapping windows, pitch 60
    centres = np.array([[60 * i, 60 * j] for i in g for j in g if (i + j) % 2 == 0], dtype=np.float64)
    assert len(centres) > 2000
    plan = ops.plan_windows(centres, S, F)
    assert plan["ok"].all()
    stamps, idx = ops.extract(field, plan, S, C, out_dtype=torch.float32)
    assert idx == list(range(len(centres)))
    # extraction window start for even F: int(F/2)-int(S/2)+d ; subtraction offset is one less
    x0 = plan["sx"]
    y0 = plan["sy"]
    res = ops.window_axpy(field, stamps, x0, y0, -1.0)
    mask = torch.zeros((F, F), dtype=torch.bool, device="cuda")
    for a, b in zip(x0[:50], y0[:50]):
        assert float(res[0, a : a + S, b : b + S].abs().max()) == 0.0
    for a, b in zip(x0, y0):
        mask[a : a + S, b : b + S] = True
    assert float(res[0][mask].abs().max()) == 0.0
    assert torch.equal(res[0][~mask], field[0][~mask])
    # scatter-add the stamps back: exact reconstruction of the windows
    back = ops.window_axpy(res, stamps, x0, y0, 1.0)
    assert torch.equal(back, field)
