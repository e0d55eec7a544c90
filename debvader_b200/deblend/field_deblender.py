"""B200-native drop-in for reference deblend/field_deblender.py (class DeblendField).

Same constructor, methods, attributes and record layout as the reference; the field lives on the
device, extraction / network / centre-MSE / subtract-back are CUDA kernels, and only the record
materialisation (a numpy recarray of per-stamp arrays, as the reference returns) touches the host.
Fractional positions in get_residual_field / get_predicted_field go through the cubic-spline placement kernels
(the reference's ndimage.shift, evaluated only where it matters); ``optimise_positions=True`` runs the
reference's own scipy least-squares call with the objective evaluated on the device
(deblend_cutout/optimization.py).
"""
import numpy as np
import pandas as pd
import torch

from .. import _fieldops
from ..deblend_cutout.deblender import deblend
from ..model.model import Deblender


class DeblendField:
    def __init__(self, net, field_image, cutout_size=59, nb_of_bands=6, epistemic_uncertainty_estimation=False, normalise=False):
        """field_deblender.py:13-44."""
        self.net = net
        if isinstance(field_image, torch.Tensor):
            self.field_image = field_image.detach().cpu().numpy().copy()
        else:
            self.field_image = np.asarray(field_image).copy()
        self.field_size = self.field_image.shape[1]
        self.cutout_size = cutout_size
        self.nb_of_bands = nb_of_bands
        self.epistemic_uncertainty_estimation = epistemic_uncertainty_estimation
        self.normalise = normalise
        self.nb_of_detected_objects = []
        self.nb_of_deblended_galaxies = []
        self.res_deblend = None
        self.mse = []
        device = net.device if isinstance(net, Deblender) else None
        self._field_dev = _fieldops.to_device_field(field_image if isinstance(field_image, torch.Tensor) else self.field_image, device)
        self._dev_cache = None  # (id(records), means_dev, stddev_dev) of the last deblend_field call

    # ------------------------------------------------------------------------------------------
    def _positions(self, res_deblend):
        """x_pos / y_pos of field_deblender.py:83-90 (float64) and whether every one is integer-valued."""
        dx = np.array([r["galaxy_distances_to_center_x"] for r in res_deblend], dtype=np.float64)
        dy = np.array([r["galaxy_distances_to_center_y"] for r in res_deblend], dtype=np.float64)
        sh = np.array([np.asarray(r["shifts"], dtype=np.float64) for r in res_deblend]).reshape(-1, 2)
        px, ix = _fieldops.positions(dx, sh[:, 0])
        py, iy = _fieldops.positions(dy, sh[:, 1])
        return px, py, ix and iy

    def _paste(self, base, stamps, px, py, integer, alpha, shape=None):
        """base + alpha * sum of the placed stamps: window copy for integer positions, cubic-spline
        ndimage.shift placement (field_deblender.py:92-95) as soon as one position is fractional."""
        if integer:
            off = _fieldops.subtract_offset(self.field_size, self.cutout_size)
            x0 = off + px.astype(np.int64)
            y0 = off + py.astype(np.int64)
            return _fieldops.window_axpy(base, stamps, x0, y0, alpha, field_shape=shape, dtype=torch.float64)
        return _fieldops.spline_window_axpy(base, stamps, px, py, alpha, field_shape=shape, dtype=torch.float64)

    def _stamps_dev(self, res_deblend, column):
        c = self._dev_cache
        if c is not None and c[0] is res_deblend and column in c[1]:
            return c[1][column]
        a = np.stack([np.asarray(r[column], dtype=np.float32) for r in res_deblend])
        return torch.from_numpy(a).to(self._field_dev.device)

    def get_residual_field(self, res_deblend=None, as_tensor=False):
        """field_deblender.py:46-97: field minus every predicted galaxy (all rows, whatever passed_cuts)."""
        if res_deblend is None:
            res_deblend = self.res_deblend
        base = self._field_dev
        if res_deblend is None or len(res_deblend) == 0:
            out = base.clone()
        else:
            px, py, integer = self._positions(res_deblend)
            out = self._paste(base, self._stamps_dev(res_deblend, "output_images_mean"), px, py, integer, -1.0)
        return out if as_tensor else out.cpu().numpy()

    def get_predicted_field(self, res_deblend=None, as_tensor=False):
        """field_deblender.py:99-189: sums of the predicted mean / stddev / epistemic stamps."""
        if res_deblend is None:
            res_deblend = self.res_deblend
        F_, C = self.field_size, self.nb_of_bands
        dev = self._field_dev.device
        names = ("predicted_mean_field", "predicted_stddev_field", "predicted_epistemic_field")
        cols = ("output_images_mean", "output_images_stddev", "epistemic_uncertainty")
        out = {}
        for name, col in zip(names, cols):
            if res_deblend is None or len(res_deblend) == 0 or (col == "epistemic_uncertainty" and not self.epistemic_uncertainty_estimation):
                f = torch.zeros((F_, F_, C), device=dev, dtype=torch.float64)
            else:
                px, py, integer = self._positions(res_deblend)
                f = self._paste(None, self._stamps_dev(res_deblend, col), px, py, integer, 1.0, shape=(F_, F_, C))
            out[name] = f if as_tensor else f.cpu().numpy()
        return out

    def get_deblending_meta_data(self, res_deblend=None):
        """field_deblender.py:191-217."""
        meta = {"field_image": self.field_image, "deblended_image": self.get_residual_field(res_deblend)}
        meta.update(self.get_predicted_field(res_deblend))
        return meta

    # ------------------------------------------------------------------------------------------
    def deblend_field(self, galaxy_distances_to_center, cutout_images=None, optimise_positions=False, epistemic_criterion=100.0,
                      mse_criterion=100.0, field_image=None):
        """field_deblender.py:219-382."""
        res_deblend = {"cutout_images": None, "output_images_mean": None, "output_images_stddev": None, "shifts": None, "list_idx": None}
        if field_image is None:
            field_dev = self._field_dev
        else:
            field_dev = _fieldops.to_device_field(field_image, self._field_dev.device)
        field_size = field_dev.shape[1]
        dev = field_dev.device
        S, C = self.cutout_size, self.nb_of_bands

        if isinstance(cutout_images, np.ndarray):
            cut_dev = torch.from_numpy(np.ascontiguousarray(cutout_images)).to(dev)
            list_idx = list(range(len(cutout_images)))
            sel = cut_dev
        else:
            plan = _fieldops.plan_windows(galaxy_distances_to_center, S, field_size)
            cut_dev, list_idx = _fieldops.extract(field_dev, plan, S, C, out_dtype=torch.float64)
            if len(list_idx) != len(plan["ok"]):
                print("Some galaxies are too close from the border of the field to be considered here.")
            sel = cut_dev[torch.as_tensor(list_idx, device=dev, dtype=torch.long)] if len(list_idx) != cut_dev.shape[0] else cut_dev
        if list_idx == []:
            print("No galaxy deblended. End of the iterative procedure.")
            return res_deblend

        # network on the device-resident stamps (deblend(): cast to fp32, net, mean / stddev)
        if isinstance(self.net, Deblender) and not self.normalise:
            dist = self.net(sel)
            mean_dev, std_dev = dist.mean().tensor, dist.stddev().tensor
        else:
            mean_np, dist = deblend(self.net, sel.cpu().numpy(), normalise=self.normalise)
            mean_dev = torch.from_numpy(np.ascontiguousarray(mean_np, dtype=np.float32)).to(dev)
            std_dev = torch.as_tensor(np.asarray(dist.stddev().numpy(), dtype=np.float32)).to(dev)

        n = len(list_idx)
        if self.epistemic_uncertainty_estimation:
            # field_deblender.py:303-316: std over 100 stochastic passes of each stamp, normalised by the r-band flux
            if isinstance(self.net, Deblender) and not self.normalise:
                e_dev = self.net.epistemic_std(sel, 100)  # batched: encoder once, 100 latent draws + decoder passes per stamp
                norm_dev = e_dev[:, :, :, 2].sum(dim=(1, 2)) / mean_dev[:, :, :, 2].double().sum(dim=(1, 2))
                epistemic = list(e_dev.cpu().numpy())
                epistemic_norm = norm_dev.cpu().numpy()
            else:
                epistemic, norm = [], []
                for i in range(n):
                    rep = sel[i : i + 1].expand(100, S, S, C).contiguous()
                    m100 = torch.as_tensor(deblend(self.net, rep.cpu().numpy(), normalise=self.normalise)[0]).to(dev)
                    e = m100.double().std(dim=0, unbiased=False)
                    epistemic.append(e.cpu().numpy())
                    norm.append(float(e[:, :, 2].sum() / mean_dev[i, :, :, 2].double().sum()))
                epistemic_norm = np.array(norm)
        else:
            epistemic = list(np.zeros((n, S, S, C)))
            epistemic_norm = np.zeros(n)

        lo, hi = int(S / 2) - 5, int(S / 2) + 5
        mse_center = _fieldops.center_mse(sel.contiguous(), mean_dev.contiguous(), lo, hi).cpu().numpy()
        passed_cuts = [not ((epistemic_norm[i] > epistemic_criterion) or (mse_center[i] > mse_criterion)) for i in range(n)]

        gx = [galaxy_distances_to_center[k][0] for k in list_idx]
        gy = [galaxy_distances_to_center[k][1] for k in list_idx]
        if optimise_positions:
            # field_deblender.py:337-352: bounded least-squares fit of a sub-pixel shift per galaxy on the r band
            # (the reference pads with self.field_size; it only works when field_image has that size too)
            from ..deblend_cutout.optimization import FieldBand, fit_position

            fb = FieldBand(field_dev)
            r_band = mean_dev[:, :, :, 2].contiguous()
            shifts = [np.array(fit_position(fb, r_band[i], galaxy_distances_to_center[k])) for i, k in enumerate(list_idx)]
        else:
            shifts = [np.array([0, 0]) for _ in range(n)]

        self.nb_of_detected_objects += [len(list(galaxy_distances_to_center))]
        self.nb_of_deblended_galaxies += [len(list_idx)]

        res_deblend["cutout_images"] = list(sel.cpu().numpy())
        res_deblend["output_images_mean"] = list(mean_dev.cpu().numpy())
        res_deblend["output_images_stddev"] = list(std_dev.cpu().numpy())
        res_deblend["shifts"] = shifts
        res_deblend["list_idx"] = list_idx
        res_deblend["galaxy_distances_to_center_x"] = gx
        res_deblend["galaxy_distances_to_center_y"] = gy
        res_deblend["epistemic_uncertainty"] = epistemic
        res_deblend["passed_cuts"] = passed_cuts

        self.res_deblend = pd.DataFrame(res_deblend).to_records(index=False)
        self._dev_cache = (self.res_deblend, {"output_images_mean": mean_dev, "output_images_stddev": std_dev})
        return self.res_deblend
