"""DeblendField / IterativeDeblendField host logic (record layout, order contract, cuts, iteration
control) against the reference's own deblend_field run with a fake net (tests/golden).  The device
operators are replaced by oracle-backed stand-ins (tests/cpu_ops.py) so this runs without a GPU."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import cpu_ops  # noqa: E402
from golden.make_golden import fake_net  # noqa: E402  (pure numpy; does not touch /root/reference on import)


@pytest.fixture()
def field_mod(monkeypatch):
    cpu_ops.install(monkeypatch)
    from debvader_b200.deblend import field_deblender

    return field_deblender


def test_deblend_field_matches_reference_records(field_mod, golden_dir):
    g = np.load(os.path.join(golden_dir, "deblend_field_fake.npz"))
    obj = field_mod.DeblendField(fake_net, g["field"], cutout_size=59, nb_of_bands=6)
    rec = obj.deblend_field(g["centres"], mse_criterion=2.0)
    assert tuple(rec.dtype.names) == tuple(g["record_names"])
    assert list(rec["list_idx"]) == list(g["list_idx"])
    assert list(rec["passed_cuts"]) == list(g["passed_cuts"])
    np.testing.assert_array_equal(np.stack(list(rec["cutout_images"])), g["cutouts"])
    np.testing.assert_array_equal(np.stack(list(rec["output_images_mean"])), g["mean"])
    np.testing.assert_array_equal(np.stack(list(rec["output_images_stddev"])), g["stddev"])
    np.testing.assert_array_equal(np.array(list(rec["galaxy_distances_to_center_x"])), g["dx"])
    np.testing.assert_array_equal(np.array(list(rec["galaxy_distances_to_center_y"])), g["dy"])
    assert obj.nb_of_detected_objects == list(g["nb_detected"])
    assert obj.nb_of_deblended_galaxies == list(g["nb_deblended"])
    np.testing.assert_allclose(obj.get_residual_field(), g["residual"], rtol=0, atol=1e-11)
    meta = obj.get_deblending_meta_data()
    assert set(meta) == {"field_image", "deblended_image", "predicted_mean_field", "predicted_stddev_field", "predicted_epistemic_field"}
    assert meta["predicted_mean_field"].shape == (101, 101, 6)


def test_no_valid_stamp_returns_dict_of_none(field_mod):
    field = np.random.default_rng(0).normal(size=(1, 40, 40, 6))
    obj = field_mod.DeblendField(fake_net, field)
    res = obj.deblend_field(np.array([[0.0, 0.0]]))  # 59-px stamp cannot fit a 40-px field
    assert isinstance(res, dict) and res["list_idx"] is None and res["cutout_images"] is None
    assert obj.res_deblend is None
    np.testing.assert_array_equal(obj.get_residual_field(), field)


def test_precomputed_cutouts_branch(field_mod, golden_dir):
    g = np.load(os.path.join(golden_dir, "deblend_field_fake.npz"))
    obj = field_mod.DeblendField(fake_net, g["field"])
    rec = obj.deblend_field(g["centres"][list(g["list_idx"])], cutout_images=g["cutouts"])
    assert list(rec["list_idx"]) == list(range(len(g["cutouts"])))
    np.testing.assert_array_equal(np.stack(list(rec["output_images_mean"])), g["mean"])


@pytest.mark.skipif(__import__("torch").cuda.is_available(), reason="checks the no-GPU failure mode")
def test_optimise_positions_has_no_cpu_fallback(field_mod):
    """the position fit evaluates its objective on the device: without one it fails loudly"""
    obj = field_mod.DeblendField(fake_net, np.zeros((1, 80, 80, 6)))
    with pytest.raises(TypeError, match="CUDA"):
        obj.deblend_field(np.array([[0, 0]]), optimise_positions=True)


def test_fractional_shifts_go_through_the_spline_placement(field_mod, golden_dir):
    """records with fractional shifts: get_residual_field / get_predicted_field route to spline_window_axpy
    (oracle-backed here) and reproduce the reference's outputs"""
    import pandas as pd

    g = np.load(os.path.join(golden_dir, "subpixel.npz"))
    name = "win_even"
    field, means, stds, pos, sh = (g[f"{name}_{k}"] for k in ("field", "means", "stds", "pos", "shifts"))
    rows = {
        "output_images_mean": list(means),
        "output_images_stddev": list(stds),
        "epistemic_uncertainty": list(np.zeros_like(means)),
        "shifts": [s for s in sh],
        "galaxy_distances_to_center_x": list(pos[:, 0]),
        "galaxy_distances_to_center_y": list(pos[:, 1]),
    }
    rec = pd.DataFrame(rows).to_records(index=False)
    obj = field_mod.DeblendField(None, field, cutout_size=means.shape[1], nb_of_bands=field.shape[3])
    np.testing.assert_allclose(obj.get_residual_field(rec), g[f"{name}_residual"], rtol=0, atol=1e-12)
    np.testing.assert_allclose(obj.get_predicted_field(rec)["predicted_stddev_field"], g[f"{name}_pred_std"], rtol=0, atol=1e-12)


def test_iterative_loop_control(monkeypatch):
    cpu_ops.install(monkeypatch)
    from debvader_b200.deblend_iterative import iterative_deblender as it

    rng = np.random.default_rng(3)
    field = rng.normal(0, 0.1, (1, 121, 121, 6))
    calls = []
    steps = [np.array([[0.0, 0.0], [10.0, 10.0]]), np.array([[5.0, -5.0], [-20.0, 3.0], [12.0, 0.0]]), np.array([[1.0, 1.0]])]

    def detector(f):
        calls.append(np.asarray(f).copy())
        return steps[min(len(calls) - 1, len(steps) - 1)]

    obj = it.IterativeDeblendField(fake_net, field, detector=detector)
    rec = obj.iterative_deblending()
    # step 1 finds 2, step 2 finds 3 (> 2, continue), step 3 finds 1 (not > 3: the loop ends after it)
    assert len(calls) == 3
    assert obj.nb_of_deblended_galaxies == [2, 3, 1]
    assert list(rec["list_idx"]) == [0, 1, 2, 3, 4, 5]  # offset by the running count (iterative_deblender.py:145-147)
    assert len(obj.mse) == 3
    # quirk kept: each step's residual restarts from the ORIGINAL field: the field the
    # detector sees at step 3 is original - step-2 galaxies only
    np.testing.assert_array_equal(calls[0], field)
    assert not np.array_equal(calls[1], field)
