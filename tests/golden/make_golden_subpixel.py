#!/usr/bin/env python
"""Generate tests/golden/subpixel.npz by running the REFERENCE's own code (build container only).

Sub-pixel placement: DeblendField.get_residual_field / get_predicted_field with non-integer
positions (deblend/field_deblender.py:46-189, scipy.ndimage.shift of a padded canvas) and
position_optimization (deblend_cutout/optimization.py:6-52, scipy.optimize.least_squares).
Reads /root/reference (read-only) and is never run on the GPU box; its output is committed.
scipy here is 1.18.1 (the reference pins 1.11.2; same ndimage spline code since 1.6).
"""
import importlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from make_golden import _stub_modules  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def blobs(rng, n, S, C):
    yy, xx = np.mgrid[0:S, 0:S].astype(np.float64)
    out = np.zeros((n, S, S, C))
    for k in range(n):
        cx, cy = (S - 1) / 2 + rng.normal(0, 1.5, 2)
        sx, sy = rng.uniform(1.5, 5.0, 2)
        g = np.exp(-0.5 * (((xx - cx) / sx) ** 2 + ((yy - cy) / sy) ** 2))
        for c in range(C):
            out[k, :, :, c] = g * 10 ** rng.uniform(-0.5, 1.2) + 0.02 * np.abs(rng.normal(size=(S, S)))
    return out.astype(np.float32)


def main():
    _stub_modules()
    fd = importlib.import_module("debvader.deblend.field_deblender")
    opt = importlib.import_module("debvader.deblend_cutout.optimization")
    import pandas as pd

    arrays = {}
    cases = {
        # name: (F, S, C, distances, shifts)
        "win_odd": (131, 59, 2, [[0.0, 0.0], [10.25, -7.5], [-30.7, 31.2], [33.0, 12.625], [60.4, -58.9], [-12.0, 4.0]],
                    [[0.3, -0.4], [0, 0], [1.1, 2.9], [0, 0], [0, 0], [0, 0]]),
        "win_even": (130, 59, 2, [[3.5, -3.5], [-20.125, 20.875], [0.0, 0.0]], [[0, 0], [0.0625, -2.0], [-3.0, 3.0]]),
        "whole": (90, 59, 2, [[0.75, -1.25], [14.5, -15.0], [-9.999, 22.3]], [[0, 0], [0.2, 0.2], [0, 0]]),
    }
    for name, (F, S, C, pos, sh) in cases.items():
        rng = np.random.default_rng(abs(hash(name)) % 1000 if False else {"win_odd": 11, "win_even": 12, "whole": 13}[name])
        field = rng.normal(0, 0.6, (1, F, F, C))
        pos = np.array(pos, dtype=np.float64)
        sh = np.array(sh, dtype=np.float64)
        means = blobs(rng, len(pos), S, C)
        stds = (0.05 * means + 1e-4).astype(np.float32)
        rows = {
            "output_images_mean": list(means),
            "output_images_stddev": list(stds),
            "epistemic_uncertainty": list(np.zeros_like(means)),
            "shifts": [s for s in sh],
            "galaxy_distances_to_center_x": list(pos[:, 0]),
            "galaxy_distances_to_center_y": list(pos[:, 1]),
        }
        rec = pd.DataFrame(rows).to_records(index=False)
        obj = fd.DeblendField(None, field, cutout_size=S, nb_of_bands=C)
        arrays[f"{name}_field"] = field
        arrays[f"{name}_pos"] = pos
        arrays[f"{name}_shifts"] = sh
        arrays[f"{name}_means"] = means
        arrays[f"{name}_stds"] = stds
        arrays[f"{name}_residual"] = obj.get_residual_field(rec)
        pf = obj.get_predicted_field(rec)
        arrays[f"{name}_pred_mean"] = pf["predicted_mean_field"]
        arrays[f"{name}_pred_std"] = pf["predicted_stddev_field"]

    # position_optimization: a field made of two shifted blobs + noise, 3 bands (band 2 is the one fitted)
    F, S, C = 131, 59, 3
    rng = np.random.default_rng(21)
    means = blobs(rng, 3, S, C)
    import scipy.ndimage

    dist = np.array([[5.0, -8.0], [-20.0, 14.0], [22.0, 25.0]])
    true_shift = np.array([[0.8, -1.3], [-0.45, 0.0], [2.2, 1.7]])
    field = rng.normal(0, 0.05, (1, F, F, C))
    off = int((F - S) / 2)
    for k in range(3):
        canvas = np.zeros((F, F, C))
        canvas[off : off + S, off : off + S] = means[k]
        for c in range(C):
            field[0, :, :, c] += scipy.ndimage.shift(canvas[:, :, c], dist[k] + true_shift[k])
    fitted = []
    for k in range(3):
        canvas = np.zeros((F, F, C))
        canvas[off : off + S, off : off + S] = means[k]
        fitted.append(opt.position_optimization(field[0], canvas, dist[k]))
    arrays["opt_field"] = field
    arrays["opt_means"] = means
    arrays["opt_dist"] = dist
    arrays["opt_true_shift"] = true_shift
    arrays["opt_fitted"] = np.array(fitted)
    print("fitted shifts", np.array(fitted))

    # deblend_field(optimise_positions=True) driven by the fake net of make_golden.py (same field / centres as
    # deblend_field_fake.npz, regenerated from the seed)
    from make_golden import fake_net

    F, S, C = 101, 59, 6
    rng = np.random.default_rng(7)
    field = rng.normal(0, 0.5, (1, F, F, C))
    field[0, 40:60, 45:58, :] += 30.0
    centres = np.array([[0.0, 0.0], [10.0, -12.0], [40.0, 0.0], [-21.0, 21.0], [21.0, 21.0], [-22.0, 0.0], [5.0, 5.0]])
    obj = fd.DeblendField(fake_net, field, cutout_size=S, nb_of_bands=C)
    rec = obj.deblend_field(centres, optimise_positions=True, mse_criterion=2.0)
    arrays["dfo_list_idx"] = np.array(list(rec["list_idx"]), dtype=np.int64)
    arrays["dfo_shifts"] = np.stack(list(rec["shifts"]))
    arrays["dfo_residual"] = obj.get_residual_field()
    print("deblend_field shifts", arrays["dfo_shifts"])
    np.savez_compressed(os.path.join(HERE, "subpixel.npz"), **arrays)
    print("written", os.path.getsize(os.path.join(HERE, "subpixel.npz")))


if __name__ == "__main__":
    main()
