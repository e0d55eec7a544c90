"""B200-native drop-in for reference extract/extraction.py:4-43."""
import numpy as np
import torch

from .. import _fieldops


def extract_cutouts(field_image, field_size, galaxy_distances_to_center, cutout_size=59, nb_of_bands=6):
    """Extract the cutouts around particular galaxies in the field.

    Same contract as the reference: returns ``(cutout_images, list_idx)`` where
    ``cutout_images`` is (N, cutout_size, cutout_size, nb_of_bands) float64 with zeros for the
    centres whose window does not fit, and ``list_idx`` the accepted indices in input order.
    The window arithmetic is planned on the host exactly as extraction.py:26-32 does it; the copy
    itself is a coalesced gather kernel (dbv_extract).  A host ndarray field is uploaded, a CUDA
    tensor field is used in place and the cutouts are then returned as a CUDA tensor.
    """
    plan = _fieldops.plan_windows(galaxy_distances_to_center, cutout_size, field_size)
    on_device = isinstance(field_image, torch.Tensor) and field_image.is_cuda
    field_dev = _fieldops.to_device_field(field_image)
    cut, list_idx = _fieldops.extract(field_dev, plan, cutout_size, nb_of_bands, out_dtype=torch.float64)
    if len(list_idx) != len(plan["ok"]):
        print("Some galaxies are too close from the border of the field to be considered here.")
    if on_device:
        return cut, list_idx
    return cut.cpu().numpy(), list_idx
