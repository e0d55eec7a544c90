"""CPU: the detection oracle (oracle/detect_numpy.py) against independent computations and hand-made cases.

`sep` is not installable here and the reference holds no golden detections: parity with sep itself is UNPINNED.  What is checked:
every stage against an independent implementation (scipy / numpy one-liners), SExtractor's documented behaviour on constructed
inputs (background of pure noise, 8-connectivity, minarea, Lutz completion order, barycentres), and the host-side wiring."""
import os

import numpy as np
import pytest
from scipy import ndimage

from debvader_b200.detect import detection as det
from oracle import detect_numpy as D


def make_field(F, n_src, seed=0, sigma=0.03, gradient=0.0, C=6):
    """noise + round Gaussian blobs (+ an optional background ramp) on every band; returns (field (1,F,F,C) f64, true (x, y))"""
    rng = np.random.default_rng(seed)
    field = rng.normal(0.0, sigma, (1, F, F, C))
    yy, xx = np.mgrid[0:F, 0:F]
    field[0] += (gradient * (xx + 2 * yy) / F)[..., None]
    pos = rng.uniform(20, F - 20, (n_src, 2))
    for (x, y) in pos:
        amp, s = rng.uniform(0.5, 3.0), rng.uniform(1.2, 2.5)
        x0, x1, y0, y1 = int(x) - 15, int(x) + 16, int(y) - 15, int(y) + 16
        blob = amp * np.exp(-((xx[y0:y1, x0:x1] - x) ** 2 + (yy[y0:y1, x0:x1] - y) ** 2) / (2 * s * s))
        field[0, y0:y1, x0:x1, :] += blob[..., None]
    return field, pos


def test_background_of_pure_noise():
    rng = np.random.default_rng(1)
    img = rng.normal(0.7, 0.2, (300, 260)).astype(np.float32)  # partial meshes on both axes
    b0, s0 = D.mesh_statistics(img)
    assert b0.shape == (5, 5) and (b0 > -D.BIG).all()
    assert np.abs(b0 - 0.7)[:4, :4].max() < 0.03 and np.abs(b0 - 0.7).max() < 0.08  # the last column of meshes is 4 px wide
    assert np.abs(s0 / 0.2 - 1)[:4, :4].max() < 0.12
    b, s, gb, gr = D.filter_meshes(b0, s0)
    assert abs(gb - 0.7) < 0.01 and abs(gr / 0.2 - 1) < 0.05
    bm = D.background_map(b, 300, 260)
    assert bm.dtype == np.float32 and bm.shape == (300, 260) and np.abs(bm - 0.7).max() < 0.05


def test_background_follows_a_ramp_and_ignores_sources():
    field, _ = make_field(512, 60, seed=2, gradient=0.05)
    img = field[0, :, :, 2].astype(np.float32)
    b, s, gb, gr = D.filter_meshes(*D.mesh_statistics(img))
    bm = D.background_map(b, 512, 512)
    yy, xx = np.mgrid[0:512, 0:512]
    assert np.abs(bm - 0.05 * (xx + 2 * yy) / 512)[96:-96, 96:-96].max() < 0.006  # the 3x3 median flattens a ramp on the outer ring of meshes
    assert abs(gr / 0.03 - 1) < 0.1


def test_spline_through_the_mesh_centres():
    """the bicubic spline interpolates: at a mesh centre (pixel 64 k + 31.5) the map equals the mesh value"""
    rng = np.random.default_rng(3)
    b = rng.normal(0, 1, (4, 5)).astype(np.float32)
    bm = D.background_map(b, 256, 320)
    mid = 0.5 * (bm[31::64][:, 31::64].astype(np.float64) + bm[32::64][:, 32::64]) * 0.5 + 0.25 * (bm[31::64][:, 32::64].astype(np.float64) + bm[32::64][:, 31::64])
    assert np.abs(mid - b).max() < 2e-3  # the four pixels around a centre average to the node value up to curvature
    # against scipy's natural cubic spline, column by column then row by row
    from scipy.interpolate import CubicSpline

    u = (np.arange(256) + 0.5) / 64 - 0.5
    node = CubicSpline(np.arange(4), b.astype(np.float64), axis=0, bc_type="natural", extrapolate=True)(u)
    v = (np.arange(320) + 0.5) / 64 - 0.5
    ref = CubicSpline(np.arange(5), node, axis=1, bc_type="natural", extrapolate=True)(v)
    # inside the outer mesh centres the two are the same spline; outside, this restatement continues the end CUBIC (as SExtractor does)
    assert np.abs(bm - ref)[32:-32, 32:-32].max() < 1e-5


def test_matched_filter_against_scipy():
    rng = np.random.default_rng(4)
    img = rng.normal(0, 1, (70, 90)).astype(np.float32)
    taps = D.normalised_filter(det.FILTER_KERNEL)
    assert abs(float(taps.sum()) - 1.0) < 1e-6 and np.array_equal(taps, det.normalised_taps())
    got = D.matched_filter(img, taps)
    want = ndimage.correlate(img.astype(np.float64), taps.astype(np.float64), mode="constant", cval=0.0)
    assert np.abs(got - want).max() < 1e-5


def _detect_mask(mask, value=1.0):
    """run the extraction stages of the oracle on a hand-made binary footprint (flat zero background, no filter)"""
    F = mask.shape[0]
    field = np.zeros((1, F, F, 6))
    field[0, :, :, 2] = mask * value
    return field


def test_connectivity_minarea_and_completion_order(monkeypatch):
    # bypass the background (a hand-made mask has no noise to estimate it from) and the filter (identity)
    monkeypatch.setattr(D, "mesh_statistics", lambda img: (np.zeros((1, 1), np.float32), np.full((1, 1), 0.1, np.float32)))
    ident = np.zeros((7, 7))
    ident[3, 3] = 1.0
    m = np.zeros((40, 40))
    # A: a "U" whose two arms end on row 12; B: a 2x2 block nested between the arms, also ending on row 12
    m[5, 5:16] = 1
    m[5:13, 5] = 1
    m[5:13, 15] = 1
    m[11:13, 9:11] = 1
    # C: a diagonal chain (8-connected) of 4 pixels; D: 3 pixels (below minarea)
    for k in range(4):
        m[20 + k, 20 + k] = 1
    m[30, 5:8] = 1
    # E: ends on row 12 as well, to the right of A
    m[10:13, 30:32] = 1
    c, dt = D.detect(_detect_mask(m), ident, return_details=True)
    assert list(dt["npix"]) == [4, 11 + 7 + 7, 6, 4]  # B (ends at x=10), A (x=15), E (x=31), then C; D dropped
    assert list(dt["last"]) == [12 * 40 + 10, 12 * 40 + 15, 12 * 40 + 31, 23 * 40 + 23]
    np.testing.assert_allclose(dt["x"][0], 9.5)
    np.testing.assert_allclose(dt["y"][0], 11.5)
    np.testing.assert_allclose([dt["x"][3], dt["y"][3]], [21.5, 21.5])
    # centres: (row, col) offsets from int(F/2), rounded half to even (np.round, detection.py:48-54)
    assert c[0].tolist() == [np.round(11.5 - 20), np.round(9.5 - 20)] == [-8.0, -10.0]


def test_blobs_are_found_where_they_are():
    field, pos = make_field(384, 40, seed=5)
    c, dt = D.detect(field, det.FILTER_KERNEL, return_details=True)
    found = np.stack([dt["x"], dt["y"]], 1)
    # every isolated true source has a detection within 0.7 px
    d = np.sqrt(((pos[:, None, :] - found[None]) ** 2).sum(-1))
    iso = np.sqrt(((pos[:, None, :] - pos[None]) ** 2).sum(-1)) + np.eye(len(pos)) * 1e9
    lone = iso.min(1) > 14
    assert lone.sum() > 20 and (d.min(1)[lone] < 0.7).all()
    assert len(c) <= len(pos) + 3  # at 1.5 sigma on the FILTERED image with minarea 4, pure-noise detections are rare
    assert (np.diff(dt["last"]) > 0).all()
    assert np.array_equal(c, np.stack([np.round(dt["y"] - 192), np.round(dt["x"] - 192)], 1))


def test_golden_detections_of_the_packaged_field(golden_dir):
    """tests/golden/detect_dc2.npz (made by tests/golden/make_golden_detect.py from THIS oracle) keeps the oracle from
    drifting; it pins nothing about sep."""
    g = np.load(os.path.join(golden_dir, "detect_dc2.npz"))
    field = np.load(os.path.join(golden_dir, "dc2_field2.npz"))["field"]
    c, dt = D.detect(field, det.FILTER_KERNEL, return_details=True)
    np.testing.assert_array_equal(c, g["centres"])
    np.testing.assert_array_equal(dt["last"], g["last"])
    assert float(dt["globalrms"]) == float(g["globalrms"])
    # most catalogue galaxies of the field are detected within 2 px (the rest are blended or below the threshold)
    truth = np.load(os.path.join(golden_dir, "dc2_field2.npz"))["centres"]
    dist = np.abs(truth[:, None, :] - c[None]).max(-1).min(1)
    assert (dist <= 2).mean() > 0.6


def test_detect_objects_wiring():
    with pytest.raises(ValueError):
        det.detect_objects(np.zeros((1, 8, 8, 6)), backend="nope")
    try:
        import sep  # noqa: F401
    except ImportError:
        with pytest.raises(ImportError, match="backend='device'"):
            det.detect_objects(np.zeros((1, 8, 8, 6)))
    d = det.DeviceDetector(max_objects=16)
    assert d.accepts_tensor and d.taps.dtype == np.float32 and d.taps.shape == (7, 7)
    from debvader_b200 import _ffi

    assert _ffi.lib().dbv_detect_scratch_bytes(4096, 4096, 1 << 17) < 1 << 30
    assert _ffi.lib().dbv_detect_scratch_bytes(0, 4096, 16) == 0


def test_against_sep_when_its_detections_are_in_the_golden(golden_dir):
    """tests/golden/make_golden_detect.py --sep (run where the reference's `sep` is installable) stores sep's own detections of the
    packaged field next to the oracle's.  Until then parity with sep is UNPINNED and this test is skipped.  What can match: the
    isolated objects and their order (this restatement has no multi-threshold deblending / clean pass, sep's gatherup is randomised)."""
    g = np.load(os.path.join(golden_dir, "detect_dc2.npz"))
    if "sep_centres" not in g.files:
        pytest.skip("no sep detections in tests/golden/detect_dc2.npz (sep not installable where the golden was made): parity unpinned")
    ours, theirs = g["centres"], g["sep_centres"]
    d = np.abs(ours[:, None, :] - theirs[None]).max(-1)
    matched = d.min(1) <= 1
    assert matched.mean() > 0.8
    # the matched objects appear in the same relative order in both lists
    j = d.argmin(1)[matched]
    assert (np.diff(j) > 0).mean() > 0.9
