"""The one polynomial-module function on the propagation path: the weighted mode sum used as the
incoherent polychromatic sum (reference prysm/polynomials/fitting.py:7-37)."""
import warnings

import numpy as np
import torch

from . import _ops


def sum_of_2d_modes(modes, weights):
    """tensordot(modes(k, m, n), weights(k)) over the leading axis as one streaming kernel."""
    if isinstance(modes, (list, tuple)):
        warnings.warn('sum_of_2d_modes: modes is a list or tuple: for optimal performance, pre convert to array of shape (k, m, n)')
        modes = torch.stack([_ops.asdevice(m) for m in modes])
    modes = _ops.asdevice(modes)
    if isinstance(weights, torch.Tensor):
        weights = weights.detach().cpu().numpy()
    return _ops.weighted_sum(modes, np.asarray(weights, dtype=np.float64))
