"""PSF reductions on the path (reference prysm/psf.py:174-237)."""
from . import _ops


def centroid(data, dx=None, unit='spatial'):
    """Centre of mass; 'pixels' is corner indexed, 'spatial' = dx*(com - shape//2)
    (prysm/psf.py:174-203; scipy.ndimage.center_of_mass = first moments / sum)."""
    if not hasattr(data, 'ndim'):
        dx = data.dx if dx is None else dx
        data = data.data
    data = _ops.asdevice(data)
    s0, sy, sx = _ops.moments(data)
    com = (sy / s0, sx / s0)
    if unit != 'spatial':
        return com
    return tuple(dx * (c - s // 2) for c, s in zip(com, data.shape))


def autocrop(data, px):
    """Window of full width px around the centroid (prysm/psf.py:206-237); a view when it fits."""
    import torch
    data = _ops.asdevice(data)
    cy, cx = (int(c) for c in centroid(data, unit='pixels'))
    w = px // 2
    y0, x0 = cy - w, cx - w
    y1, x1 = y0 + px, x0 + px
    pad_y = (max(0, -y0), max(0, y1 - data.shape[0]))
    pad_x = (max(0, -x0), max(0, x1 - data.shape[1]))
    if any(pad_y) or any(pad_x):
        data = torch.nn.functional.pad(data, (pad_x[0], pad_x[1], pad_y[0], pad_y[1]))
        y0, y1, x0, x1 = y0 + pad_y[0], y1 + pad_y[0], x0 + pad_x[0], x1 + pad_x[0]
    return data[y0:y1, x0:x1]
