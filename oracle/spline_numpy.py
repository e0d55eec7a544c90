"""numpy restatement of the sub-pixel placement of the reference (oracle; test infrastructure).

The reference places a predicted stamp at a non-integer position with
``scipy.ndimage.shift(canvas, shift=(x_pos, y_pos))`` on a zero canvas of the size of the field
(deblend/field_deblender.py:66-95 for the residual, :121-182 for the predicted fields,
deblend_cutout/optimization.py:27-29,41-44 for the position fit).  The arithmetic lives in a
third-party dependency that is not under /root/reference: **scipy==1.11.2** (requirements.txt:7;
the build container has 1.18.1, same algorithm since 1.6).  Defaults of that call: order=3,
mode='constant', cval=0.0, prefilter=True.  Restated here from scipy's published algorithm
(ndimage/src/ni_splines.c, ni_interpolation.c:NI_ZoomShift):

* ``spline_filter1d_mirror``  cubic B-spline prefilter of one axis: gain (1-z)(1-1/z), pole
  z = sqrt(3)-2, causal + anti-causal recursion, *mirror* boundary initialisation (scipy uses the
  mirror initialisation for mode='constant').
* ``shift_cubic_constant``    out[i,j] = sum of 4x4 taps of the prefiltered canvas around
  (i-sx, j-sy), B-spline weights of ``get_spline_interpolation_weights``; an output whose source
  coordinate lies outside [0, n-1] is cval=0; taps outside [0, n-1] are mirror-mapped.
* ``residual_field_subpixel`` / ``predicted_field_subpixel``  the reference loops with that shift.
* ``position_objective``      ``fun`` of optimization.py:21-33.

PINNED: tests/test_oracle_golden.py checks ``shift_cubic_constant`` against scipy.ndimage.shift
itself and ``residual_field_subpixel`` against the output of the reference's own
``DeblendField.get_residual_field`` with non-integer positions (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np

POLE = np.sqrt(3.0) - 2.0


def spline_filter1d_mirror(a, axis):
    """ni_splines.c: apply_filter for order 3 (one pole), mirror boundaries, along `axis`."""
    z = POLE
    c = np.moveaxis(np.array(a, dtype=np.float64, copy=True), axis, 0)
    n = c.shape[0]
    if n < 2:
        return np.moveaxis(c, 0, axis)
    c *= (1.0 - z) * (1.0 - 1.0 / z)
    # _init_causal_mirror (exact sum over the whole line)
    z_n_1 = z ** (n - 1)
    c0 = c[0] + z_n_1 * c[n - 1]
    z_i = z
    for i in range(1, n - 1):
        c0 = c0 + z_i * (c[i] + z_n_1 * c[n - 1 - i])
        z_i *= z
    c[0] = c0 / (1.0 - z_n_1 * z_n_1)
    for i in range(1, n):
        c[i] += z * c[i - 1]
    # _init_anticausal_mirror
    c[n - 1] = (z * c[n - 2] + c[n - 1]) * z / (z * z - 1.0)
    for i in range(n - 2, -1, -1):
        c[i] = z * (c[i + 1] - c[i])
    return np.moveaxis(c, 0, axis)


def spline_weights(x):
    """ni_interpolation.c: get_spline_interpolation_weights(x, order=3) -> (start index, 4 weights)."""
    fl = np.floor(x)
    y = x - fl
    z = 1.0 - y
    w1 = (y * y * (y - 2.0) * 3.0 + 4.0) / 6.0
    w2 = (z * z * (z - 2.0) * 3.0 + 4.0) / 6.0
    w0 = z * z * z / 6.0
    w3 = 1.0 - w0 - w1 - w2
    return fl.astype(np.int64) - 1, np.stack([w0, w1, w2, w3], axis=-1)


def _mirror_index(idx, n):
    """tap index mapping of NI_ZoomShift for every mode but grid-constant."""
    if n <= 1:
        return np.zeros_like(idx)
    s2 = 2 * n - 2
    idx = np.array(idx, dtype=np.int64, copy=True)
    neg = idx < 0
    t = idx[neg]
    t = s2 * (-t // s2) + t
    t = np.where(t <= 1 - n, t + s2, -t)
    idx[neg] = t
    big = idx >= n
    t = idx[big]
    t = t - s2 * (t // s2)
    t = np.where(t >= n, s2 - t, t)
    idx[big] = t
    return idx


def _axis_taps(n, shift):
    """per output index: valid flag, 4 (mirror-mapped) tap indices, 4 weights."""
    cc = np.arange(n, dtype=np.float64) + (-float(shift))  # zoom_shift passes shift = -shift
    valid = ~((cc < 0) | (cc > n - 1))
    start, w = spline_weights(np.where(valid, cc, 0.0))
    taps = _mirror_index(start[:, None] + np.arange(4)[None, :], n)
    return valid, taps, w


def shift_cubic_constant(img, shift):
    """scipy.ndimage.shift(img, shift) for a 2-D array with the defaults the reference uses."""
    img = np.asarray(img, dtype=np.float64)
    coef = spline_filter1d_mirror(spline_filter1d_mirror(img, 0), 1)
    n0, n1 = img.shape
    v0, t0, w0 = _axis_taps(n0, shift[0])
    v1, t1, w1 = _axis_taps(n1, shift[1])
    out = np.zeros_like(img)
    # NI_ZoomShift accumulates the taps with axis 0 outermost
    for a in range(4):
        rows = coef[t0[:, a]] * w0[:, a][:, None]
        for b in range(4):
            out += rows[:, t1[:, b]] * w1[:, b][None, :]
    out[~v0, :] = 0.0
    out[:, ~v1] = 0.0
    return out


def _padded(stamp_band, field_size, cutout_size):
    off = int((field_size - cutout_size) / 2)  # field_deblender.py:72
    canvas = np.zeros((field_size, field_size))
    canvas[off : cutout_size + off, off : cutout_size + off] = stamp_band
    return canvas


def residual_field_subpixel(field_image, means, pos_x, pos_y, cutout_size=59):
    """get_residual_field (field_deblender.py:46-97) literally: one full-canvas spline shift per
    galaxy and band, subtracted in row order.  O(N * C * F^2): small cases only."""
    out = np.array(field_image, dtype=np.float64, copy=True)
    F_ = out.shape[1]
    for m, px, py in zip(means, pos_x, pos_y):
        for band in range(out.shape[3]):
            out[0, :, :, band] -= shift_cubic_constant(_padded(m[:, :, band], F_, cutout_size), (px, py))
    return out


def predicted_field_subpixel(field_size, nb_of_bands, stamps, pos_x, pos_y, cutout_size=59):
    """one of the three accumulators of get_predicted_field (field_deblender.py:99-189)."""
    acc = np.zeros((field_size, field_size, nb_of_bands))
    for m, px, py in zip(stamps, pos_x, pos_y):
        for band in range(nb_of_bands):
            acc[:, :, band] += shift_cubic_constant(_padded(m[:, :, band], field_size, cutout_size), (px, py))
    return acc


def position_objective(x, r_band_field, net_output):
    """``fun`` of deblend_cutout/optimization.py:21-33."""
    return np.square(r_band_field - shift_cubic_constant(net_output, (x[0], x[1]))).mean()
