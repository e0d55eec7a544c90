"""Distribution-like return objects.

The reference returns TensorFlow-Probability distributions from ``net(x)`` /
``decoder(z)`` (tfd.Normal, model/model.py:154-159) and ``z(x)``
(tfd.MultivariateNormalTriL, model/model.py:211-214); callers use ``.mean()``,
``.stddev()``, ``.sample(n)``, ``.log_prob(x)`` and ``.numpy()`` on the results
(deblend_cutout/deblender.py:24, deblend/field_deblender.py:368-370,
notebooks/behavior_of_latent_space.ipynb).  These classes give the same surface
over torch tensors that stay on the device until ``.numpy()`` is called.
"""
from __future__ import annotations

import math

import numpy as np
import torch


class Value:
    """A tensor result with the ``.numpy()`` accessor TF eager tensors have."""

    __slots__ = ("tensor",)

    def __init__(self, t):
        self.tensor = t if isinstance(t, torch.Tensor) else torch.as_tensor(t)

    def numpy(self):
        return self.tensor.detach().cpu().numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.numpy()
        return a.astype(dtype) if dtype is not None else a

    @property
    def shape(self):
        return tuple(self.tensor.shape)

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, i):
        return Value(self.tensor[i])

    def __repr__(self):
        return f"Value(shape={self.shape}, device={self.tensor.device})"


class NormalOutput:
    """Independent Normal(loc, scale) per pixel — the decoder's DistributionLambda."""

    def __init__(self, loc, scale, generator=None):
        self._loc = loc if isinstance(loc, torch.Tensor) else torch.as_tensor(loc)
        self._scale = None if scale is None else (scale if isinstance(scale, torch.Tensor) else torch.as_tensor(scale))
        self._gen = generator

    def mean(self):
        return Value(self._loc)

    def stddev(self):
        if self._scale is None:
            raise RuntimeError("this result was computed without the stddev output")
        return Value(self._scale)

    def sample(self, n=None):
        shape = self._loc.shape if n is None else (int(n),) + tuple(self._loc.shape)
        eps = torch.randn(shape, device=self._loc.device, dtype=self._loc.dtype, generator=self._gen)
        return Value(self._loc + self.stddev().tensor * eps)

    def log_prob(self, x):
        x = torch.as_tensor(np.asarray(x) if not isinstance(x, torch.Tensor) else x).to(self._loc.device, self._loc.dtype)
        s = self.stddev().tensor
        return Value(-0.5 * ((x - self._loc) / s) ** 2 - torch.log(s) - 0.5 * math.log(2 * math.pi))


class MVNTriLOutput:
    """MultivariateNormalTriL(loc, scale_tril) of the latent layer: only what callers use."""

    def __init__(self, loc, stddev, sample=None):
        self._loc, self._std, self._sample = loc, stddev, sample

    def mean(self):
        return Value(self._loc)

    def stddev(self):
        return Value(self._std)

    def sample(self):
        if self._sample is None:
            raise RuntimeError("no sample was drawn for this result")
        return Value(self._sample)
