"""numpy restatement of the field-level hot path (oracle; test infrastructure).

* ``extract_cutouts``      follows extract/extraction.py:4-43
* ``plan_windows``         the index arithmetic of extraction.py:26-32 made explicit
* ``residual_field``       slice form of DeblendField.get_residual_field,
                           deblend/field_deblender.py:46-97 (integer positions)
* ``predicted_fields``     slice form of get_predicted_field, field_deblender.py:99-189
* ``center_mse``           field_deblender.py:323-332 (+ training/metrics.py:4-12)
* ``mse``                  training/metrics.py:4-12

PINNED: tests/test_oracle_golden.py checks these against outputs of the
reference's own code (tests/golden/make_golden.py imports it from
/root/reference in the build container).
"""
from __future__ import annotations

import numpy as np


def window_start(shift, cutout_size: int, field_size: int) -> int:
    """extraction.py:26 / :29 — note int() truncates toward zero."""
    return -int(cutout_size / 2) + int(shift) + int(field_size / 2)


def plan_windows(galaxy_distances_to_center, cutout_size: int, field_size: int):
    """For each centre return (start_x, len_x, start_y, len_y, accepted).

    numpy basic-slicing + assignment-broadcast semantics of
    ``cutout_images[i] = field_image[0, xs:xe, ys:ye]`` (extraction.py:32):
    the slice is clipped / wrapped by ``slice.indices``; the assignment succeeds
    iff each axis has length S or 1 (a length-1 axis broadcasts); otherwise numpy
    raises ValueError and the reference skips the stamp (extraction.py:35-36).
    """
    S, F_ = cutout_size, field_size
    out = []
    for c in galaxy_distances_to_center:
        try:
            xs = window_start(c[0], S, F_)
            ys = window_start(c[1], S, F_)
        except ValueError:  # int(nan)
            out.append((0, 0, 0, 0, False))
            continue
        xe = xs + 2 * int(S / 2) + 1
        ye = ys + 2 * int(S / 2) + 1
        rx = range(*slice(xs, xe).indices(F_))
        ry = range(*slice(ys, ye).indices(F_))
        ok = len(rx) in (S, 1) and len(ry) in (S, 1)
        out.append((rx.start if len(rx) else 0, len(rx), ry.start if len(ry) else 0, len(ry), ok))
    return out


def extract_cutouts(field_image, field_size, galaxy_distances_to_center, cutout_size=59, nb_of_bands=6):
    """extraction.py:4-43 restated through plan_windows (no try/except on the copy)."""
    n = len(galaxy_distances_to_center)
    S = cutout_size
    cut = np.zeros((n, S, S, nb_of_bands))
    idx = []
    Cf = field_image.shape[-1]
    chan_ok = Cf == nb_of_bands or Cf == 1
    for i, (sx, lx, sy, ly, ok) in enumerate(plan_windows(galaxy_distances_to_center, S, field_size)):
        if not (ok and chan_ok):
            continue
        cut[i] = field_image[0, sx : sx + lx, sy : sy + ly]
        idx.append(i)
    return cut, idx


def subtract_offset(field_size: int, cutout_size: int) -> int:
    """field_deblender.py:72 — pos_offset = int((F - S) / 2)."""
    return int((field_size - cutout_size) / 2)


def _paste(acc, stamp, x0, y0, sign):
    """acc[x0:x0+S, y0:y0+S] += sign*stamp, clipped to the field: scipy.ndimage.shift
    with mode='constant' drops whatever leaves the canvas (field_deblender.py:92-95)."""
    FH, FW = acc.shape[0], acc.shape[1]  # square for a whole field; a rank's local region of a tiled field is rectangular
    S = stamp.shape[0]
    ax0, ax1 = max(x0, 0), min(x0 + S, FH)
    ay0, ay1 = max(y0, 0), min(y0 + S, FW)
    if ax0 >= ax1 or ay0 >= ay1:
        return
    part = stamp[ax0 - x0 : ax1 - x0, ay0 - y0 : ay1 - y0].astype(np.float64)
    if sign < 0:
        acc[ax0:ax1, ay0:ay1] -= part
    else:
        acc[ax0:ax1, ay0:ay1] += part


def residual_field(field_image, means, pos_x, pos_y, cutout_size=59):
    """get_residual_field (field_deblender.py:46-97) for integer x_pos / y_pos.

    field_image (1,F,F,C) f64; means (N,S,S,C) f32; pos = distance + shift.
    Stamps are subtracted one after the other in row order (fp64, the
    float32 mean widened exactly), which is what the kernel must reproduce.
    """
    out = field_image.copy()
    F_ = field_image.shape[1]
    off = subtract_offset(F_, cutout_size)
    for m, px, py in zip(means, pos_x, pos_y):
        _paste(out[0], m, off + int(px), off + int(py), -1)
    return out


def predicted_fields(field_size, nb_of_bands, means, stddevs, epistemic, pos_x, pos_y, cutout_size=59):
    """get_predicted_field (field_deblender.py:99-189) for integer positions."""
    acc = [np.zeros((field_size, field_size, nb_of_bands)) for _ in range(3)]
    off = subtract_offset(field_size, cutout_size)
    for i, (px, py) in enumerate(zip(pos_x, pos_y)):
        for a, src in zip(acc, (means, stddevs, epistemic)):
            if src is not None:
                _paste(a, src[i], off + int(px), off + int(py), +1)
    return {"predicted_mean_field": acc[0], "predicted_stddev_field": acc[1], "predicted_epistemic_field": acc[2]}


def mse(a, b):
    """training/metrics.py:4-12."""
    return np.mean(np.square(a - b))


def center_mse(cutouts, means, cutout_size=59):
    """field_deblender.py:323-332: window [int(S/2)-5, int(S/2)+5) on both axes, all bands."""
    lo = int(cutout_size / 2) - 5
    hi = int(cutout_size / 2) + 5
    return np.array([mse(c[lo:hi, lo:hi], m[lo:hi, lo:hi]) for c, m in zip(cutouts, means)])
