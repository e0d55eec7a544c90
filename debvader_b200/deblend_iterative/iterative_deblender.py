"""B200-native drop-in for reference deblend_iterative/iterative_deblender.py.

Loop control is the reference's, statement for statement, including its quirks:
``get_residual_field()`` inside the loop always starts from the ORIGINAL field and only the
current step's records; ``epistemic_criterion`` is not forwarded after the first step.  One
deliberate fix: a step that deblends nothing ends the iteration instead of raising TypeError
on ``len(None)`` (iterative_deblender.py:141).
"""
import numpy as np

from ..deblend.field_deblender import DeblendField
from ..detect.detection import detect_objects
from ..training.metrics import mse


class IterativeDeblendField(DeblendField):
    def __init__(self, net, field_image, cutout_size=59, nb_of_bands=6, epistemic_uncertainty_estimation=False, normalise=False,
                 detector=None):
        super().__init__(net, field_image, cutout_size, nb_of_bands, epistemic_uncertainty_estimation, normalise)
        self.detector = detector or detect_objects  # extension: any callable field -> (N,2) centres

    def iterative_deblending(self, galaxy_distances_to_center=None, cutout_images=None, optimise_positions=False,
                             epistemic_criterion=100.0, mse_criterion=100.0):
        """iterative_deblender.py:21-99."""
        field_image = self.field_image.copy()
        res_step = self.deblending_step(field_image, cutout_images=cutout_images, optimise_positions=optimise_positions,
                                        epistemic_criterion=epistemic_criterion, mse_criterion=mse_criterion)
        res_deblend = res_step
        if res_step is None or res_step["list_idx"] is None:
            print("converged !")
            self.res_deblend = None
            return self.res_deblend

        new_residual_field = self.get_residual_field()
        self.mse += [mse(self.field_image, new_residual_field)]
        shifts_previous = []
        k = 1
        diff_mse = -1

        while len(res_step["shifts"]) > len(shifts_previous):
            print(f"iteration {k}")
            shifts_previous = res_step["shifts"]
            prev_residual_field = new_residual_field
            res_step = self.deblending_step(prev_residual_field, cutout_images=None, optimise_positions=optimise_positions,
                                            mse_criterion=mse_criterion)
            if res_step is None or res_step["list_idx"] is None:
                break
            new_residual_field = self.get_residual_field()
            self.mse += [mse(prev_residual_field, new_residual_field)]
            res_deblend = np.concatenate([res_deblend, res_step])
            k += 1
            print(f"{sum(self.nb_of_deblended_galaxies)} galaxies found up to this step.")
            print(f"deta_mse = {diff_mse}, mse_iteration = " + str(self.mse[-1]) + " and mse_previous_step = " + str(self.mse[-2]))

        print("converged !")
        self.res_deblend = res_deblend
        self._dev_cache = None
        return self.res_deblend

    def deblending_step(self, field_image, cutout_images=None, optimise_positions=False, epistemic_criterion=100.0, mse_criterion=100.0):
        """iterative_deblender.py:101-152."""
        detection_k = self.detector(field_image)
        res_step = self.deblend_field(field_image=field_image, galaxy_distances_to_center=detection_k, cutout_images=cutout_images,
                                      optimise_positions=optimise_positions, epistemic_criterion=epistemic_criterion,
                                      mse_criterion=mse_criterion)
        if res_step["list_idx"] is None or len(res_step["list_idx"]) == 0:
            print("No more galaxies found")
            return res_step if isinstance(res_step, dict) else None
        res_step["list_idx"] += sum(self.nb_of_deblended_galaxies) - self.nb_of_deblended_galaxies[-1]
        print(f"Deblend {self.nb_of_deblended_galaxies[-1]} more galaxy(ies)")
        return res_step
