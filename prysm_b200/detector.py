"""Detector sampling at the end of the image chain (reference prysm/detector.py:151-338): binning to detector pixels, its
adjoint, and the analytic pixel / optical-low-pass-filter transfer functions that `apply_transfer_functions` consumes.
`Detector.expose` (random shot / read noise) is not on the deterministic path and is not provided."""
import numbers

from . import _ops


def _factors(factor, ndim=2):
    if isinstance(factor, numbers.Number):
        return (int(factor),) * ndim
    f = tuple(int(v) for v in factor)
    if len(f) != ndim:
        raise ValueError(f'factor must have {ndim} entries for a {ndim}-D array')
    return f


def bindown(array, factor, mode='avg'):
    """Bin a 2-D array by `factor` (int or (fy, fx)); the shape must be an integer multiple (prysm/detector.py:222-274)."""
    if mode.lower() in ('avg', 'average', 'mean'):
        mean = True
    elif mode.lower() == 'sum':
        mean = False
    else:
        raise ValueError('mode must be average or sum.')
    fy, fx = _factors(factor)
    return _ops.bindown(array, fy, fx, mean)


def tile(array, factor, scaling='sum'):
    """Repeat every sample `factor` times per axis -- the adjoint of bindown (prysm/detector.py:277-338)."""
    fy, fx = _factors(factor)
    if scaling == 'sum':
        sf = 1 / (fy * fx)
    elif scaling in ('avg', 'average', 'mean'):
        sf = 1
    else:
        raise ValueError('scaling must be average or sum')
    return _ops.tile(array, fy, fx, sf)


def pixel_ft(fx, fy, width_x, width_y):
    """sinc(fx wx) sinc(fy wy) on broadcastable frequency vectors fx (1, N), fy (M, 1) (prysm/detector.py:174-194)."""
    return _ops.separable_tf(0, fx, fy, width_x, width_y)


def olpf_ft(fx, fy, width_x, width_y):
    """cos(2 wx fx) cos(2 wy fy) (prysm/detector.py:151-171)."""
    return _ops.separable_tf(1, fx, fy, width_x, width_y)
