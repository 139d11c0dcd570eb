"""Circular aperture masks on the device (reference prysm/geometry.py:11-34, 337-372)."""
from . import _ops


def circle_sdf(radius, r):
    """Signed distance to a circle, negative inside (prysm/geometry.py:337-353)."""
    return _ops.asdevice(r) - radius


def circle(radius, r):
    """Binary mask r - radius <= 0 as a bool tensor (prysm/geometry.py:356-372)."""
    return _ops.circle(_ops.asdevice(r), radius)


def antialias(d, dx):
    """Signed distance -> coverage with a one-sample edge ramp, clip(0.5 - d/dx, 0, 1) (prysm/geometry.py:11-34)."""
    d = _ops.asdevice(d)
    return _ops.circle(d, 0.0, aa_dx=dx)          # the circle kernel with radius 0 ramps any signed distance


def grey_circle(radius, r, dx):
    """antialias(circle_sdf(radius, r), dx) in one pass: the grey-edge aperture coronagraph models need
    (prysm/propagation/coronagraph.py:107-110)."""
    return _ops.circle(_ops.asdevice(r), radius, aa_dx=dx)
