"""B200-native drop-in for reference deblend_cutout/optimization.py (position_optimization).

The reference minimises, with scipy.optimize.least_squares (bounds +-3 px, 2-point Jacobian), the scalar
``fun(x) = mean((r_band_field - ndimage.shift(net_output, x))**2)`` where ``net_output`` is the padded
r-band prediction already shifted to the detected position — every evaluation is a cubic-spline shift of
the WHOLE field-sized canvas (optimization.py:21-46).  Here the optimiser is the same scipy call (host
control flow, a dependency the reference already has), and every evaluation of ``fun`` runs on the
device on the only pixels where the shifted prediction is not negligible
(``dbv_position_objective``: two placed windows + one fixed-order reduction; the field's own sum of
squares is computed once).  fp64, same arithmetic as scipy's spline code to ~1e-15 relative, so the
optimiser follows the same path up to the noise of its own finite differences.
"""
import ctypes as C

import numpy as np
import torch
from scipy import optimize

from .. import _ffi, _fieldops

R_BAND = 2  # optimization.py:35-36


class FieldBand:
    """One band of a device-resident (1,F,F,C) float64 field + its sum of squares (computed once)."""

    def __init__(self, field_dev, band=R_BAND):
        _fieldops._require_cuda(field_dev, "field")  # no CPU path
        if field_dev.dtype != torch.float64:
            field_dev = field_dev.double()
        self.field = field_dev.contiguous()
        self.F, self.C, self.band = int(self.field.shape[-3]), int(self.field.shape[-1]), int(band)
        lib = _ffi.lib()
        sb = int(lib.dbv_mse_scratch_bytes())
        scratch = torch.empty((sb // 8,), device=self.field.device, dtype=torch.float64)
        out = torch.empty((1,), device=self.field.device, dtype=torch.float64)
        with torch.cuda.device(self.field.device):
            _ffi.check(lib.dbv_band_sumsq(_ffi.ptr(self.field), self.F, self.C, self.band, _ffi.ptr(out), _ffi.ptr(scratch), sb, _ffi.stream_ptr()))
        self.sumsq = float(out.item())


def fit_position(fb: FieldBand, stamp_band_dev, galaxy_distance_to_center, margin=_fieldops.SPLINE_MARGIN, return_result=False):
    """position_optimization for one predicted stamp: stamp_band_dev (S,S) CUDA (the r band of the prediction)."""
    dev = fb.field.device
    S = int(stamp_band_dev.shape[0])
    d = np.asarray(galaxy_distance_to_center, dtype=np.float64)
    placed1, a1x, a1y = _fieldops.spline_place(stamp_band_dev.reshape(1, S, S, 1).contiguous(), d[0:1], d[1:2], fb.F, margin)
    E1 = int(placed1.shape[-1])
    E2 = _fieldops.spline_extent(E1, margin)
    scratch = torch.empty((int(_ffi.lib().dbv_spline_scratch_doubles(1, E1, 1, int(margin))),), device=dev, dtype=torch.float64)
    placed2 = torch.empty((E2 * E2,), device=dev, dtype=torch.float64)
    out_dev = torch.empty((1,), device=dev, dtype=torch.float64)
    out_host = C.c_double(0.0)
    lib = _ffi.lib()
    args = (_ffi.ptr(fb.field), fb.F, fb.C, fb.band, _ffi.ptr(placed1), E1, int(a1x[0]), int(a1y[0]))
    tail = (int(margin), fb.sumsq, _ffi.ptr(scratch), _ffi.ptr(placed2), _ffi.ptr(out_dev), C.cast(C.byref(out_host), C.c_void_p))

    def fun(x):
        with torch.cuda.device(dev):
            _ffi.check(lib.dbv_position_objective(*args, float(x[0]), float(x[1]), *tail, _ffi.stream_ptr()))
        return out_host.value

    opt = optimize.least_squares(fun, (0.0, 0.0), bounds=(-3, 3))  # optimization.py:37-49
    if return_result:
        return opt
    return opt.x[0], opt.x[1]


def fit_positions(field_dev, r_band_batch, centres, margin=_fieldops.SPLINE_MARGIN):
    """position_optimization for all galaxies of a field: r_band_batch (N,S,S) CUDA, centres (N,2) -> (N,2) shifts."""
    fb = FieldBand(field_dev)
    return np.array([fit_position(fb, r_band_batch[i].contiguous(), centres[i], margin) for i in range(len(centres))], dtype=np.float64).reshape(-1, 2)


def position_optimization(field_image, output_image_mean_padded, galaxy_distance_to_center, cutout_size=59):
    """optimization.py:6-52, same arguments (field_image (F,F,C), the padded prediction (F,F,C), the detected
    offset); `cutout_size` (extension) tells where the stamp sits inside the padded canvas."""
    field = np.asarray(field_image, dtype=np.float64)
    F = field.shape[0]
    dev_field = _fieldops.to_device_field(field[None])
    fb = FieldBand(dev_field)
    off = _fieldops.subtract_offset(F, cutout_size)
    block = np.ascontiguousarray(np.asarray(output_image_mean_padded)[off : off + cutout_size, off : off + cutout_size, R_BAND], dtype=np.float64)
    return fit_position(fb, torch.from_numpy(block).to(dev_field.device), galaxy_distance_to_center)
