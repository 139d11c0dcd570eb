"""The polynomial-module functions on the propagation path: the weighted mode sum used as the incoherent
polychromatic sum and as the modal OPD synthesis, and its adjoint (reference prysm/polynomials/fitting.py:7-57)."""
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


def sum_of_2d_modes_adjoint(modes, databar):
    """tensordot(modes(k, m, n), databar(m, n)) over the trailing axes: k reductions in one launch; the (k,)
    gradient comes back as a host array like the weights it pairs with (prysm/polynomials/fitting.py:40-57)."""
    if isinstance(modes, (list, tuple)):
        modes = torch.stack([_ops.asdevice(m) for m in modes])
    modes = _ops.asdevice(modes)
    if modes.dtype not in (torch.float32, torch.float64):
        modes = modes.to(torch.float64)
    return _ops.mode_projection(modes, _ops.asdevice(databar))
