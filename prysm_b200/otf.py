"""PSF -> MTF / PTF / OTF on the B200 engine (reference prysm/otf.py:11-202).

One fused centred FFT of the real PSF (shifts folded into the passes, no complex copy of the
input) and one normalisation kernel that divides by the centre sample and takes abs / angle.
"""
from . import _ops
from ._richdata import RichData


def _unwrap_psf(psf, dx):
    """prysm/otf.py:16-25."""
    if not hasattr(psf, 'ndim'):
        dx = psf.dx
        psf = psf.data
    if dx is None:
        raise ValueError('dx is None: dx must be provided if psf is an array')
    return _ops.asdevice(psf), dx


def transform_psf(psf, dx=None):
    """fftshift(fft2(ifftshift(psf))), unnormalised; df = 1000/(rows*dx) (prysm/otf.py:28-33)."""
    psf, dx = _unwrap_psf(psf, dx)
    data = _ops.fft2(psf, tuple(psf.shape), dir=-1, scale=1.0, shift_in=True, shift_out=True)
    return data, 1000 / (data.shape[0] * dx)


def transform_psf_adjoint(data_bar):
    """fftshift(ifft2(ifftshift(g), norm='forward')) (prysm/otf.py:36-59)."""
    g = _ops.ascomplex(_ops.asdevice(data_bar))
    return _ops.fft2(g, tuple(g.shape), dir=+1, scale=1.0, shift_in=True, shift_out=True)


def _normalized(psf, dx, which):
    data, df = transform_psf(psf, dx)
    mtf, ptf, otf = _ops.otf_normalize(data, which)
    return mtf, ptf, otf, data, df


def mtf_from_psf(psf, dx=None, return_more=False):
    """prysm/otf.py:77-104."""
    mtf, _, _, data, df = _normalized(psf, dx, 1)
    rd = RichData(data=mtf, dx=df, wavelength=None)
    return (rd, data) if return_more else rd


def ptf_from_psf(psf, dx=None, return_more=False):
    """prysm/otf.py:107-137."""
    _, ptf, _, data, df = _normalized(psf, dx, 2)
    rd = RichData(data=ptf, dx=df, wavelength=None)
    return (rd, data) if return_more else rd


def otf_from_psf(psf, dx=None, return_more=False):
    """prysm/otf.py:140-167."""
    _, _, otf, data, df = _normalized(psf, dx, 4)
    rd = RichData(data=otf, dx=df, wavelength=None)
    return (rd, data) if return_more else rd


def mtf_ptf_otf_from_psf(psf, dx=None, return_more=False):
    """All three from one forward transform (prysm/otf.py:170-202)."""
    mtf, ptf, otf, data, df = _normalized(psf, dx, 7)
    out = (RichData(mtf, df, None), RichData(ptf, df, None), RichData(otf, df, None))
    return out + (data,) if return_more else out


def encircled_energy(psf, dx, radius, return_more=False):
    """Baliga & Cohn encircled energy at `radius` [um] (scalar or iterable) from the MTF of the PSF
    (prysm/otf.py:346-387): one reduction kernel per call covers all radii."""
    import numbers
    import numpy as np
    mtf, data = mtf_from_psf(psf, dx, return_more=True)
    scalar = isinstance(radius, numbers.Number)
    radii = np.atleast_1d(np.asarray(radius, dtype=np.float64)) / 1e3
    out = _ops.encircled_energy(mtf.data, mtf.dx, radii)
    out = float(out[0]) if scalar else out
    return (out, data) if return_more else out
