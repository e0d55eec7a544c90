"""B200-native drop-in for reference deblend/field_deblender.py (class DeblendField).

Same constructor, methods, attributes and record layout as the reference; the field lives on the
device, extraction / network / centre-MSE / subtract-back are CUDA kernels, and the per-stamp record
columns are lazy device-backed proxies (``_records.DeviceStamp``: an ndarray to every caller, fetched
from the device only when somebody looks at the values).  Fractional positions in get_residual_field /
get_predicted_field go through the cubic-spline placement kernels (the reference's ndimage.shift,
evaluated only where it matters); ``optimise_positions=True`` fits all galaxies at once on the device
(deblend_cutout/optimization.py).

Multi-GPU (extension): ``DeblendField(..., tiled=True)`` inside a torch.distributed job
(one process per GPU) keeps only this rank's owner tile + 30-px halo of the field on the device
(``parallel.LocalField``); ``deblend_field`` then deblends the sources this rank owns, and
``get_residual_field`` exchanges the overlapping stamps once (NCCL all_to_all) and assembles the
rank's region, bit-identical to the single-GPU result.
"""
import numpy as np
import torch

from .. import _fieldops, _records
from ..deblend_cutout.deblender import deblend
from ..model.model import Deblender

_EMPTY = {"cutout_images": None, "output_images_mean": None, "output_images_stddev": None, "shifts": None, "list_idx": None}


class DeblendField:
    def __init__(self, net, field_image, cutout_size=59, nb_of_bands=6, epistemic_uncertainty_estimation=False, normalise=False,
                 *, tiled=False, group=None):
        """field_deblender.py:13-44.  ``field_image``: (1,F,F,C) ndarray (copied, as the reference does) or CUDA tensor
        (kept on the device; ``.field_image`` downloads it on first access).  tiled=True: see the module docstring."""
        self.net = net
        self.cutout_size = cutout_size
        self.nb_of_bands = nb_of_bands
        self.epistemic_uncertainty_estimation = epistemic_uncertainty_estimation
        self.normalise = normalise
        self.nb_of_detected_objects = []
        self.nb_of_deblended_galaxies = []
        self.res_deblend = None
        self.mse = []
        device = net.device if isinstance(net, Deblender) else None
        self._group = group
        self._local = None
        self._tile_state = None  # (records, TilePlan, mine) of the last tiled deblend_field call
        self._host_field = None
        if tiled:
            from .. import parallel

            rank, world = parallel._world(group)
            if isinstance(field_image, parallel.LocalField):
                self._local = field_image
            else:
                self._local = parallel.LocalField.from_full(field_image, rank, world, device)
                if not isinstance(field_image, torch.Tensor):
                    self._host_field = field_image  # the caller's full host array (not copied: every rank would hold 8 copies)
            self._field_dev = self._local.data
            self.field_size = self._local.field_size
        else:
            if not isinstance(field_image, torch.Tensor):
                self._host_field = np.asarray(field_image).copy()
            self._field_dev = _fieldops.to_device_field(field_image if isinstance(field_image, torch.Tensor) else self._host_field, device)
            self.field_size = self._field_dev.shape[1]

    @property
    def field_image(self):
        """the reference's ``self.field_image`` (host ndarray); tiled fields gather their owner tiles."""
        if self._host_field is None:
            if self._local is not None:
                from .. import parallel

                self._host_field = parallel.gather_field(self._local, group=self._group)
            else:
                self._host_field = self._field_dev.detach().cpu().numpy()
        return self._host_field

    @property
    def field_tensor(self):
        """the device-resident field (tiled: this rank's local region)."""
        return self._field_dev

    # ------------------------------------------------------------------------------------------
    def _positions(self, res_deblend):
        """x_pos / y_pos of field_deblender.py:83-90 (float64) and whether every one is integer-valued."""
        pc = getattr(self, "_pos_cache", None)
        if pc is not None and pc[0] is res_deblend:
            px, ix = _fieldops.positions(pc[1], 0.0)
            py, iy = _fieldops.positions(pc[2], 0.0)
            return px, py, ix and iy
        dx = np.asarray(res_deblend["galaxy_distances_to_center_x"], dtype=np.float64)
        dy = np.asarray(res_deblend["galaxy_distances_to_center_y"], dtype=np.float64)
        sh = np.array([np.asarray(s, dtype=np.float64) for s in res_deblend["shifts"]]).reshape(-1, 2)
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
        return _records.column_tensor(res_deblend, column, self._field_dev.device)

    def _tiled_paste(self, res_deblend, column, alpha, base):
        """tiled get_residual_field / get_predicted_field: exchange the overlapping stamps of `column` once and apply
        them to this rank's region in ascending global index."""
        from .. import parallel

        st = self._tile_state
        if st is None or st[0] is not res_deblend:
            raise NotImplementedError("a tiled DeblendField assembles the records of its own last deblend_field call")
        _, tp, mine, xp = st
        px, py, integer = self._positions(res_deblend) if len(res_deblend) else (None, None, True)
        if not integer:
            raise NotImplementedError("tiled fields place stamps on whole pixels only")
        S, C = self.cutout_size, self.nb_of_bands
        dev = self._field_dev.device
        own = self._stamps_dev(res_deblend, column) if len(res_deblend) else torch.empty((0, S, S, C), device=dev, dtype=torch.float32)
        stamps, ids = parallel.exchange_halo_stamps(own.contiguous(), mine, tp.owner, tp.touches, self._group, plan=xp)
        if len(ids) == 0:
            return self._field_dev.clone() if base == "field" else torch.zeros_like(self._field_dev)
        return parallel.subtract_local(self._local, tp, stamps, ids, alpha, base=base, plan=xp)

    def get_residual_field(self, res_deblend=None, as_tensor=False):
        """field_deblender.py:46-97: field minus every predicted galaxy (all rows, whatever passed_cuts).
        as_tensor=True keeps the result on the device (tiled: this rank's local region, halo included)."""
        if res_deblend is None:
            res_deblend = self.res_deblend
        base = self._field_dev
        if self._local is not None:
            from .. import parallel

            if res_deblend is None:
                res_deblend = self._tile_state[0] if self._tile_state is not None else None
            out = base.clone() if res_deblend is None else self._tiled_paste(res_deblend, "output_images_mean", -1.0, "field")
            return out if as_tensor else parallel.gather_field(self._local, out, self._group)
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
            skip = col == "epistemic_uncertainty" and not self.epistemic_uncertainty_estimation
            if self._local is not None:
                from .. import parallel

                if res_deblend is None and self._tile_state is not None:
                    res_deblend = self._tile_state[0]
                f = torch.zeros_like(self._field_dev) if (res_deblend is None or skip) else self._tiled_paste(res_deblend, col, 1.0, "zeros")
                out[name] = f if as_tensor else parallel.gather_field(self._local, f.to(torch.float64), self._group)[0]
                continue
            if res_deblend is None or len(res_deblend) == 0 or skip:
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

    def field_mse(self, a, b):
        """training/metrics.py:4-12 of two fields on the device (tiled: owner-tile partial sums + one all-reduce)."""
        if self._local is not None:
            from .. import parallel

            return parallel.field_mse_tiled(self._local, a, b, self._group)
        return _fieldops.mse(a, b)

    # ------------------------------------------------------------------------------------------
    @_records.with_quiet_gc
    def deblend_field(self, galaxy_distances_to_center, cutout_images=None, optimise_positions=False, epistemic_criterion=100.0,
                      mse_criterion=100.0, field_image=None):
        """field_deblender.py:219-382.  Tiled: the records are those of the sources THIS rank owns, ``list_idx`` stays
        global (the order contract holds across ranks: concatenating the ranks' records and sorting by list_idx gives
        the single-GPU records)."""
        res_deblend = dict(_EMPTY)
        S, C = self.cutout_size, self.nb_of_bands
        tp = mine = xp = None
        if self._local is not None:
            from .. import parallel

            if cutout_images is not None or optimise_positions:
                raise NotImplementedError("tiled fields: precomputed cutouts / position fits are single-GPU options")
            local = self._local
            if field_image is not None:
                local = field_image if isinstance(field_image, parallel.LocalField) else self._local.like(field_image)
            field_dev = local.data
            dev = field_dev.device
            tp = parallel.TilePlan(galaxy_distances_to_center, local.field_size, local.world, S)
            # exchange / subtraction indices go to the device now, while it is idle (see parallel.ExchangePlan)
            xp = parallel.ExchangePlan(tp, local.rank, local.world, dev, local.region)
            mine = xp.mine
            if len(tp.idx) != tp.n_sources:
                print("Some galaxies are too close from the border of the field to be considered here.")
            if len(tp.idx) == 0:
                print("No galaxy deblended. End of the iterative procedure.")
                self._tile_state = None
                return res_deblend
            sel = parallel.extract_local(local, tp, mine, C, out_dtype=torch.float64)
            list_idx = [int(i) for i in tp.idx[mine]]
            n_detected, n_deblended = tp.n_sources, len(tp.idx)
        else:
            field_dev = self._field_dev if field_image is None else _fieldops.to_device_field(field_image, self._field_dev.device)
            field_size = field_dev.shape[1]
            dev = field_dev.device
            if isinstance(cutout_images, (np.ndarray, torch.Tensor)):
                sel = cutout_images if isinstance(cutout_images, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(cutout_images))
                sel = sel.to(dev)
                list_idx = list(range(len(cutout_images)))
            else:
                plan = _fieldops.plan_windows(galaxy_distances_to_center, S, field_size)
                cut_dev, list_idx = _fieldops.extract(field_dev, plan, S, C, out_dtype=torch.float64)
                if len(list_idx) != len(plan["ok"]):
                    print("Some galaxies are too close from the border of the field to be considered here.")
                sel = cut_dev[torch.as_tensor(list_idx, device=dev, dtype=torch.long)] if len(list_idx) != cut_dev.shape[0] else cut_dev
            if list_idx == []:
                print("No galaxy deblended. End of the iterative procedure.")
                return res_deblend
            n_detected, n_deblended = len(list(galaxy_distances_to_center)), len(list_idx)

        n = len(list_idx)
        # network on the device-resident stamps (deblend(): cast to fp32, net, mean / stddev)
        if n == 0:
            mean_dev = torch.empty((0, S, S, C), device=dev, dtype=torch.float32)
            std_dev = torch.empty_like(mean_dev)
        elif isinstance(self.net, Deblender) and not self.normalise:
            dist = self.net(sel)
            mean_dev, std_dev = dist.mean().tensor, dist.stddev().tensor
        else:
            mean_np, dist = deblend(self.net, sel.cpu().numpy(), normalise=self.normalise)
            mean_dev = torch.from_numpy(np.ascontiguousarray(mean_np, dtype=np.float32)).to(dev)
            std_dev = torch.as_tensor(np.asarray(dist.stddev().numpy(), dtype=np.float32)).to(dev)

        if self.epistemic_uncertainty_estimation and n:
            # field_deblender.py:303-316: std over 100 stochastic passes of each stamp, normalised by the r-band flux
            if isinstance(self.net, Deblender) and not self.normalise:
                e_dev = self.net.epistemic_std(sel, 100)  # batched: encoder once, 100 latent draws + decoder passes per stamp
            else:
                e_dev = torch.empty((n, S, S, C), device=dev, dtype=torch.float64)
                for i in range(n):
                    rep = sel[i : i + 1].expand(100, S, S, C).contiguous()
                    m100 = torch.as_tensor(deblend(self.net, rep.cpu().numpy(), normalise=self.normalise)[0]).to(dev)
                    e_dev[i] = m100.double().std(dim=0, unbiased=False)
            epistemic_norm = (e_dev[:, :, :, 2].sum(dim=(1, 2)) / mean_dev[:, :, :, 2].double().sum(dim=(1, 2))).cpu().numpy()
            epistemic = _records.stamp_column(e_dev)
        else:
            # the reference stores n separate zero maps (field_deblender.py:317-321): one shared read-only map here
            zero = np.zeros((S, S, C))
            zero.setflags(write=False)
            epistemic = np.frompyfunc(lambda _i: zero, 1, 1)(np.arange(n))
            epistemic_norm = np.zeros(n)

        # everything above only ENQUEUED device work; the record columns are built on the host while the GPU runs, and the
        # one value the host needs from the device (the centre-window MSE behind passed_cuts) is fetched last
        lo, hi = int(S / 2) - 5, int(S / 2) + 5
        mse_dev = _fieldops.center_mse(sel.contiguous(), mean_dev.contiguous(), lo, hi) if n else None
        gdc = galaxy_distances_to_center
        if isinstance(gdc, np.ndarray) and gdc.ndim == 2:
            li = np.asarray(list_idx, dtype=np.int64)
            gx, gy = gdc[li, 0], gdc[li, 1]
        else:
            gx = [gdc[k][0] for k in list_idx]
            gy = [gdc[k][1] for k in list_idx]
        col_cut = _records.stamp_column(sel)
        col_mean = _records.stamp_column(mean_dev)
        col_std = _records.stamp_column(std_dev)
        if optimise_positions:
            # field_deblender.py:337-352: bounded least-squares fit of a sub-pixel shift per galaxy on the r band
            # (the reference pads with self.field_size; it only works when field_image has that size too)
            from ..deblend_cutout.optimization import fit_positions

            r_band = mean_dev[:, :, :, 2].contiguous()
            fitted = fit_positions(field_dev, r_band, np.array([[gdc[k][0], gdc[k][1]] for k in list_idx], dtype=np.float64))
            shifts = np.empty(n, dtype=object)
            for i in range(n):
                shifts[i] = np.array(fitted[i])
        else:
            # np.array([0, 0]) per galaxy (field_deblender.py:354): rows of one (n, 2) array, gathered by a C-level loop
            shifts = np.frompyfunc(np.zeros((n, 2), dtype=np.int64).__getitem__, 1, 1)(np.arange(n))
        self.nb_of_detected_objects += [n_detected]
        self.nb_of_deblended_galaxies += [n_deblended]

        # the records are assembled BEFORE the host waits for the device (the network is still running): only the column
        # that depends on a device result — passed_cuts, from the centre-window MSE — is filled in afterwards
        cols = {
            "cutout_images": col_cut,
            "output_images_mean": col_mean,
            "output_images_stddev": col_std,
            "shifts": shifts,
            "list_idx": np.asarray(list_idx, dtype=np.int64),
            "galaxy_distances_to_center_x": gx,
            "galaxy_distances_to_center_y": gy,
            "epistemic_uncertainty": epistemic,
            "passed_cuts": np.zeros(n, dtype=bool),
        }
        self.res_deblend = _records.make_records(cols)
        mse_center = mse_dev.cpu().numpy() if n else np.zeros(0)
        if n:
            self.res_deblend["passed_cuts"][:] = ~((epistemic_norm > epistemic_criterion) | (mse_center > mse_criterion))
        if not optimise_positions:  # integer shifts (0, 0): positions known without walking the records again
            self._pos_cache = (self.res_deblend, np.asarray(gx, dtype=np.float64), np.asarray(gy, dtype=np.float64))
        if tp is not None:
            self._tile_state = (self.res_deblend, tp, mine, xp)
        return self.res_deblend
