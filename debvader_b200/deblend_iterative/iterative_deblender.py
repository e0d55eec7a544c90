"""B200-native drop-in for reference deblend_iterative/iterative_deblender.py.

Loop control is the reference's, statement for statement, including its quirks:
``get_residual_field()`` inside the loop always starts from the ORIGINAL field and only the
current step's records; ``epistemic_criterion`` is not forwarded after the first step.  One
deliberate fix: a step that deblends nothing ends the iteration instead of raising TypeError
on ``len(None)`` (iterative_deblender.py:141).

The field never leaves the device between steps: residual fields are CUDA tensors, the field MSE is a
device reduction, and a detector that declares ``accepts_tensor = True`` is handed the tensor itself
(the reference's ``detect_objects`` wraps the CPU library ``sep`` and needs a host array: that download
is then the only full-field transfer of a step; ``detector="device"`` runs the detection kernels of
csrc/detect_kernels.cu on the device-resident field instead).  With ``tiled=True`` (one process per GPU) every rank
keeps its owner tile + halo only; a detector with ``accepts_local = True`` is called as
``detector(region_tensor, local_field)`` and must return the GLOBAL list of centres, identical on every
rank — any other detector gets the gathered host field (on every rank, so that all ranks see the same list).
"""
import numpy as np
import torch

from ..deblend.field_deblender import DeblendField
from ..detect.detection import DeviceDetector, TiledDeviceDetector, detect_objects


class IterativeDeblendField(DeblendField):
    def __init__(self, net, field_image, cutout_size=59, nb_of_bands=6, epistemic_uncertainty_estimation=False, normalise=False,
                 detector=None, *, tiled=False, group=None):
        super().__init__(net, field_image, cutout_size, nb_of_bands, epistemic_uncertainty_estimation, normalise, tiled=tiled, group=group)
        # extension: any callable field -> (N,2) centres; "device" = the CUDA detector (detect/detection.py:DeviceDetector),
        # which takes the device-resident residual field as it is: no field-sized transfer per iteration
        if isinstance(detector, str):
            if detector not in ("device", "sep"):
                raise ValueError(f"unknown detector {detector!r}")
            if detector == "device":
                detector = (TiledDeviceDetector(group=group, device=self._field_dev.device) if self._local is not None
                            else DeviceDetector(device=self._field_dev.device))
            else:
                detector = detect_objects
        self.detector = detector or detect_objects

    def iterative_deblending(self, galaxy_distances_to_center=None, cutout_images=None, optimise_positions=False,
                             epistemic_criterion=100.0, mse_criterion=100.0):
        """iterative_deblender.py:21-99."""
        field_image = self._field_dev  # the reference copies the host array; the device tensor is never written to
        res_step = self.deblending_step(field_image, cutout_images=cutout_images, optimise_positions=optimise_positions,
                                        epistemic_criterion=epistemic_criterion, mse_criterion=mse_criterion)
        res_deblend = res_step
        if res_step is None or res_step["list_idx"] is None:
            print("converged !")
            self.res_deblend = None
            return self.res_deblend

        new_residual_field = self.get_residual_field(as_tensor=True)
        self.mse += [self.field_mse(self._field_dev, new_residual_field)]
        # len(res_step["shifts"]) of iterative_deblender.py:58 = the number of galaxies the step deblended =
        # nb_of_deblended_galaxies[-1] (on a tiled field the GLOBAL count: every rank takes the same branch)
        n_step, n_previous = self.nb_of_deblended_galaxies[-1], 0
        k = 1
        diff_mse = -1

        while n_step > n_previous:
            print(f"iteration {k}")
            n_previous = n_step
            prev_residual_field = new_residual_field
            res_step = self.deblending_step(prev_residual_field, cutout_images=None, optimise_positions=optimise_positions,
                                            mse_criterion=mse_criterion)
            if res_step is None or res_step["list_idx"] is None:
                break
            n_step = self.nb_of_deblended_galaxies[-1]
            new_residual_field = self.get_residual_field(as_tensor=True)
            self.mse += [self.field_mse(prev_residual_field, new_residual_field)]
            res_deblend = np.concatenate([res_deblend, res_step])
            k += 1
            print(f"{sum(self.nb_of_deblended_galaxies)} galaxies found up to this step.")
            print(f"deta_mse = {diff_mse}, mse_iteration = " + str(self.mse[-1]) + " and mse_previous_step = " + str(self.mse[-2]))

        print("converged !")
        self.res_deblend = res_deblend
        return self.res_deblend

    def _detect(self, field_image):
        det = self.detector
        if isinstance(field_image, torch.Tensor):
            if self._local is not None and getattr(det, "accepts_local", False):
                return det(field_image, self._local)
            if self._local is None and getattr(det, "accepts_tensor", False):
                return det(field_image)
            if self._local is not None:
                from .. import parallel

                field_image = parallel.gather_field(self._local, field_image, self._group)
            else:
                field_image = field_image.detach().cpu().numpy()
        return det(field_image)

    def deblending_step(self, field_image, cutout_images=None, optimise_positions=False, epistemic_criterion=100.0, mse_criterion=100.0):
        """iterative_deblender.py:101-152."""
        detection_k = self._detect(field_image)
        res_step = self.deblend_field(field_image=field_image, galaxy_distances_to_center=detection_k, cutout_images=cutout_images,
                                      optimise_positions=optimise_positions, epistemic_criterion=epistemic_criterion,
                                      mse_criterion=mse_criterion)
        if res_step["list_idx"] is None or (self._local is None and len(res_step["list_idx"]) == 0):
            print("No more galaxies found")
            return res_step if isinstance(res_step, dict) else None
        res_step["list_idx"] += sum(self.nb_of_deblended_galaxies) - self.nb_of_deblended_galaxies[-1]
        print(f"Deblend {self.nb_of_deblended_galaxies[-1]} more galaxy(ies)")
        return res_step
