"""Minimal stand-in for prysm._richdata.RichData (reference prysm/_richdata.py:41-92): the
container `Wavefront.intensity` and the otf functions return.  Only the fields the hot path
produces are kept (data, dx, wavelength); plotting / interpolation stay with the reference.
"""
from . import _ops


class RichData:
    """Array plus its sample spacing and wavelength."""

    def __init__(self, data, dx, wavelength):
        self.data = data
        self.dx = dx
        self.wavelength = wavelength

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def support_x(self):
        return self.shape[1] * self.dx

    @property
    def support_y(self):
        return self.shape[0] * self.dx

    def numpy(self):
        """Host copy of the data."""
        return _ops.asnumpy(self.data)

    def copy(self):
        return RichData(self.data.clone(), self.dx, self.wavelength)
