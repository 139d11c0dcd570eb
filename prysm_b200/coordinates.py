"""Coordinate grids on the device (reference prysm/coordinates.py:73-102, 344-378) -- the inputs of the pupil
synthesis that precedes the propagation path."""
import torch

from . import _ops
from .conf import config


def make_xy_grid(shape, *, dx=0, diameter=0, grid=True):
    """x, y = fftrange(n)*dx per axis in config precision (prysm/coordinates.py:344-378); one kernel for both
    meshgrids.  shape is (rows, cols); diameter, if given, sets dx = diameter / max(shape)."""
    if not isinstance(shape, tuple):
        shape = (shape, shape)
    if diameter != 0:
        dx = diameter / max(shape)
    g = _ops.xy_grid(shape, dx, config.real_dtype, _ops.device(), want='xy')
    if grid:
        return g['x'], g['y']
    return g['x'][0].contiguous(), g['y'][:, 0].contiguous()


def make_polar_grid(shape, *, dx=0, diameter=0):
    """r, t of the same grid without materialising x, y (cart_to_polar(*make_xy_grid(...)) in one kernel)."""
    if not isinstance(shape, tuple):
        shape = (shape, shape)
    if diameter != 0:
        dx = diameter / max(shape)
    g = _ops.xy_grid(shape, dx, config.real_dtype, _ops.device(), want='rt')
    return g['r'], g['t']


def cart_to_polar(x, y, vec_to_grid=True):
    """rho = hypot(x, y), phi = arctan2(y, x) (prysm/coordinates.py:73-102); 1-D vectors broadcast to a grid."""
    x, y = _ops.asdevice(x), _ops.asdevice(y)
    if x.dtype not in (torch.float32, torch.float64):
        x = x.to(config.real_dtype)
    if vec_to_grid and x.ndim == 1:
        ny, nx = y.shape[0], x.shape[0]
        x, y = x[None, :].expand(ny, nx), y[:, None].expand(ny, nx)
    return _ops.cart_to_polar(x, y)
