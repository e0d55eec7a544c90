"""-m gpu: the device detector (csrc/detect_kernels.cu through dbv_detect) against oracle/detect_numpy.py,
stage by stage and BIT FOR BIT (background meshes, background map, matched filter, labels, order, barycentres, centres).
Parity with `sep` itself is unpinned (see the oracle's header)."""
import os

import numpy as np
import pytest
import torch

from debvader_b200.detect import detection as det
from oracle import detect_numpy as D
from tests.test_detect_oracle import make_field

pytestmark = pytest.mark.gpu


def stage_report(field, detector=None):
    """every stage of the device detector compared with the oracle; returns {stage: number of mismatching values}"""
    detector = detector or det.DeviceDetector()
    c_ref, o = D.detect(field, det.FILTER_KERNEL, return_details=True)
    c_dev, d = detector(torch.as_tensor(field).cuda(), return_details=True)
    rep = {}
    for name, key in (("back_raw", "back_raw"), ("sigma_raw", "sigma_raw"), ("back", "back"), ("sigma", "sigma"), ("fg", "fg"), ("conv", "conv")):
        got = detector.plane(name).cpu().numpy()
        rep[name] = int((got != o[key]).sum())
    rep["globalback"] = int(np.float32(d["globalback"]) != o["globalback"])
    rep["globalrms"] = int(np.float32(d["globalrms"]) != o["globalrms"])
    rep["thresh"] = int(np.float32(d["thresh"]) != o["thresh"])
    lab = detector.plane("label").cpu().numpy()
    mask = o["conv"] > o["thresh"]
    rep["mask"] = int(((lab >= 0) != mask).sum())
    # the device labels a component with its smallest raster index
    from scipy import ndimage

    ref_lab, n = ndimage.label(mask, structure=np.ones((3, 3), int))
    if n:
        idx = np.arange(mask.size).reshape(mask.shape)
        roots = ndimage.minimum(idx, ref_lab, index=np.arange(1, n + 1)).astype(np.int64)
        want = np.where(mask, roots[np.maximum(ref_lab, 1) - 1], -1)
        rep["labels"] = int((want != lab).sum())
    rep["n"] = int(len(c_dev) != len(c_ref))
    if len(c_dev) == len(c_ref):
        rep["npix"] = int((d["npix"] != o["npix"]).sum())
        rep["x"] = int((d["x"] != o["x"]).sum())
        rep["y"] = int((d["y"] != o["y"]).sum())
        rep["centres"] = int((c_dev != c_ref).sum())
    return rep, c_dev, c_ref


@pytest.mark.parametrize("case", ["dc2_259", "synthetic_1000_ramp", "tiny_70x70"])
def test_device_detector_matches_the_oracle_bit_for_bit(case, golden_dir):
    if case == "dc2_259":  # the packaged DC2 field: 259 = 4 meshes + 3 px, partial meshes on both axes
        field = np.load(os.path.join(golden_dir, "dc2_field2.npz"))["field"]
    elif case == "synthetic_1000_ramp":
        field, _ = make_field(1000, 400, seed=11, gradient=0.05)
    else:
        field, _ = make_field(70, 3, seed=12)
    rep, c_dev, c_ref = stage_report(field)
    assert not any(rep.values()), rep
    assert c_dev.dtype == np.float64 and c_dev.shape == c_ref.shape and len(c_dev) > 0
    if case == "dc2_259":
        np.testing.assert_array_equal(c_dev, np.load(os.path.join(golden_dir, "detect_dc2.npz"))["centres"])


def test_float32_field_and_repeatability():
    field, _ = make_field(640, 150, seed=13)
    d = det.DeviceDetector()
    t = torch.as_tensor(field).cuda()
    a = d(t)
    b = d(t)
    np.testing.assert_array_equal(a, b)  # integer atomics + fixed-order sums: run-to-run identical
    f32 = field.astype(np.float32)
    c = d(torch.as_tensor(f32).cuda())
    np.testing.assert_array_equal(c, D.detect(f32, det.FILTER_KERNEL))
    np.testing.assert_array_equal(det.detect_objects(t), a)  # CUDA tensors default to the device backend
    np.testing.assert_array_equal(det.detect_objects(field, backend="device"), a)  # host arrays are uploaded


def test_4096_field_properties():
    """BASELINE cfg 4 size: 4096^2 x 6 f64, 2000 sources: size-independent properties + the oracle's extraction stages."""
    field, pos = make_field(4096, 2000, seed=14)
    d = det.DeviceDetector()
    t = torch.as_tensor(field).cuda()
    c, info = d(t, return_details=True)
    lab = d.plane("label").cpu().numpy()
    conv = d.plane("conv").cpu().numpy()
    assert ((lab >= 0) == (conv > info["thresh"])).all()
    roots = np.unique(lab[lab >= 0])
    assert (lab.ravel()[roots] == roots).all()  # every label is a root
    # completion order: the largest raster index of the objects ascends
    flat = lab.ravel()
    order = np.argsort(flat, kind="stable")
    last_of = {}
    fl = flat[order]
    start = np.searchsorted(fl, roots)
    end = np.searchsorted(fl, roots, side="right")
    sizes = end - start
    keep = sizes >= 4
    last = np.array([order[s:e].max() for s, e in zip(start[keep], end[keep])])
    assert len(c) == keep.sum() == len(info["npix"])
    assert sorted(info["npix"].tolist()) == sorted(sizes[keep].tolist())
    np.testing.assert_array_equal(np.sort(last), last[np.argsort(last)])
    # sources: every isolated one has a detection within 0.7 px
    found = np.stack([info["x"], info["y"]], 1)
    from scipy.spatial import cKDTree

    dd, _ = cKDTree(found).query(pos)
    nn, _ = cKDTree(pos).query(pos, k=2)
    lone = nn[:, 1] > 14
    assert lone.sum() > 1000 and (dd[lone] < 0.7).mean() > 0.995
    assert abs(float(info["globalrms"]) / 0.03 - 1) < 0.05
    # E1-E4 of the oracle on the device's own foreground plane: filtered image, order, barycentres and centres bit for bit at full size
    # (the background stages are compared bit for bit on the smaller fields above: their pure-Python histogram loop takes minutes here)
    fg = d.plane("fg").cpu().numpy()
    conv_ref, objs = D.extract_objects(fg, D.normalised_filter(det.FILTER_KERNEL), np.float32(info["thresh"]))
    assert np.array_equal(conv_ref, conv)
    c_ref, x_ref, y_ref = D.centres_of(objs, 4096, 4096)
    np.testing.assert_array_equal(c, c_ref)
    np.testing.assert_array_equal(info["x"], x_ref)
    np.testing.assert_array_equal(info["y"], y_ref)


def test_iterative_deblending_with_the_device_detector(golden_dir):
    """IterativeDeblendField(detector="device"): detection -> extraction -> network -> subtract, all on the device."""
    from debvader import IterativeDeblendField
    from debvader.model.model import load_deblender
    from oracle import weights as ow

    field = np.load(os.path.join(golden_dir, "dc2_field2.npz"))["field"]
    net = load_deblender("dc2", (59, 59, 6), 32, [32, 64, 128, 256], [3, 3, 3, 3], weights=ow.make_random_weights(seed=1234), precision="bf16x3", seed=3)
    it = IterativeDeblendField(net, field, detector="device")
    first = it.detector(it._field_dev)
    np.testing.assert_array_equal(first, np.load(os.path.join(golden_dir, "detect_dc2.npz"))["centres"])
    res = it.iterative_deblending()
    assert res is not None and len(res) >= it.nb_of_deblended_galaxies[0] > 0
    # the first step deblended exactly the accepted detections, in detection order
    n0 = it.nb_of_deblended_galaxies[0]
    from oracle import field_numpy as fo

    _, idx = fo.extract_cutouts(field, 259, first, 59, 6)
    assert n0 == len(idx)
    np.testing.assert_array_equal(np.array(list(res["galaxy_distances_to_center_x"][:n0])), first[idx, 0])
    np.testing.assert_array_equal(np.array(list(res["galaxy_distances_to_center_y"][:n0])), first[idx, 1])
    net.close()


def test_tiled_detector_outside_a_distributed_job_is_the_plain_detector():
    from debvader_b200 import parallel

    field, _ = make_field(320, 40, seed=15)
    t = torch.as_tensor(field).cuda()
    local = parallel.LocalField.from_full(t, 0, 1)
    np.testing.assert_array_equal(det.TiledDeviceDetector()(local.data, local), det.DeviceDetector()(t))
    assert det.TiledDeviceDetector.meshes_covered(4096, 8, 30) and det.TiledDeviceDetector.meshes_covered(4097, 8, 30)
    assert not det.TiledDeviceDetector.meshes_covered(4097, 8, 0)  # without a halo a mesh astride a tile edge belongs to nobody


def test_two_gpu_tiled_detection_is_bit_identical():
    """2 ranks over NCCL: owner tile + halo each, mesh statistics all-reduced, objects all-gathered — the single-GPU list on every
    rank; a footprint wider than the halo takes the assembled-field path.  Skipped on a one-GPU box; the tiling logic of the
    kernels is also run on CPU by tests/test_detect_emul.py (2 and 8 emulated ranks)."""
    import json
    import subprocess
    import sys

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29547", os.path.join(root, "tools", "detect_tiled_nccl.py"), "1024", "300"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads([l for l in out.stdout.splitlines() if l.startswith("{")][-1])
    assert line["plain"]["identical_to_single_gpu_on_every_rank"] and line["plain"]["assembled_field_fallbacks"] == 0
    assert line["wide_object"]["identical_to_single_gpu_on_every_rank"] and line["wide_object"]["assembled_field_fallbacks"] > 0
    assert line["plain"]["region_share"] < 0.6
