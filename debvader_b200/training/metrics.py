"""B200-native drop-in for reference training/metrics.py."""
import numpy as np
import torch

from .. import _fieldops


def mse(img1, img2):
    """mean((img1-img2)^2) — training/metrics.py:4-12 — on the device, fp64, fixed reduction order.
    Returns a Python float."""
    return _fieldops.mse(img1, img2)


def vae_loss(ground_truth, predicted_distribution):
    """-log_prob of the ground truth under the predicted distribution — training/metrics.py:16-26."""
    return -predicted_distribution.log_prob(ground_truth).tensor
