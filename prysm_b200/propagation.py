"""Pupil <-> focus and plane-to-plane propagation on the B200 engine.

Same names, argument meaning, units and error behaviour as prysm.propagation (reference
prysm/propagation/{fft,dft,angular_spectrum,_kernels,wavefront}.py); arrays are CUDA tensors
(host arrays are uploaded).  Each array-level function is ONE libprysm_b200 call in which the
pad, both shifts, the normalisation -- and where asked the phase-screen synthesis and |.|^2 --
are fused into the FFT passes instead of being separate full-array copies.

Units: wavelength um, pupil dx mm, focal dx um, OPD nm, efl / z mm.  Index order [y, x].
"""
import math
import numbers
from collections.abc import Iterable

import numpy as np
import torch

from . import _ops
from ._capi import OUT_COMPLEX, OUT_INTENSITY, OUT_ACCUMULATE
from ._richdata import RichData
from .conf import config
from .fttools import pad2d, crop_center, MDFT, CZT, FFTDFT


# ------------------------------------------------------------------------------------------
# array API: FFT focus family (prysm/propagation/fft.py)
# ------------------------------------------------------------------------------------------

def _padded_shape(shape, Q):
    return tuple(shape) if Q == 1 else tuple(math.ceil(s * Q) for s in shape)


def _shape_before_pad(shape, Q):
    """prysm/propagation/_kernels.py:14-18."""
    return tuple(shape) if Q == 1 else tuple(int(s // Q) for s in shape)


def _field(w):
    return _ops.ascomplex(_ops.asdevice(w))


def focus(wavefunction, Q):
    """Pupil -> PSF plane: fftshift(fft2(ifftshift(pad2d(w, Q)), norm='ortho')) as one fused call
    (prysm/propagation/fft.py:7-25)."""
    w = _field(wavefunction)
    ky, kx = _padded_shape(w.shape, Q)
    return _ops.fft2(w, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True)


def unfocus(wavefunction, Q):
    """PSF -> pupil plane, the same with ifft2 (prysm/propagation/fft.py:48-65)."""
    w = _field(wavefunction)
    ky, kx = _padded_shape(w.shape, Q)
    return _ops.fft2(w, (ky, kx), dir=+1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True)


def focus_adjoint(wavefunction, Q):
    """Adjoint of focus: centred ortho ifft2 then crop_center to shape//Q
    (prysm/propagation/fft.py:28-45)."""
    w = _field(wavefunction)
    ky, kx = w.shape
    return _ops.fft2(w, (ky, kx), dir=+1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True,
                     crop=_shape_before_pad(w.shape, Q))


def unfocus_adjoint(wavefunction, Q):
    """prysm/propagation/fft.py:68-85."""
    w = _field(wavefunction)
    ky, kx = w.shape
    return _ops.fft2(w, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True,
                     crop=_shape_before_pad(w.shape, Q))


def focus_intensity(wavefunction, Q, weight=1.0, out=None):
    """|focus(w, Q)|^2 with the modulus fused into the last FFT pass (no complex field is
    written).  With `out`, accumulates out += weight*|.|^2 -- the per-wavelength term of the
    incoherent sum (prysm/propagation/wavefront.py:147-151, prysm/polynomials/fitting.py:37)."""
    w = _field(wavefunction)
    ky, kx = _padded_shape(w.shape, Q)
    kind = OUT_INTENSITY if out is None else OUT_ACCUMULATE
    return _ops.fft2(w, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True,
                     out_kind=kind, weight=weight, out=out)


def psf_from_amp_and_phase(amplitude, phase, wavelength, Q, weight=1.0, out=None, field=False):
    """from_amp_and_phase -> focus -> (intensity) in one call: the phase screen
    A*exp(i*2*pi/wvl*OPD) is synthesised inside the first FFT pass
    (prysm/propagation/wavefront.py:59-79, 478-504, 147-151)."""
    opd = _ops.asdevice(phase)
    if opd.dtype not in (torch.float32, torch.float64):
        opd = opd.to(config.real_dtype)
    amp = None if amplitude is None else _ops.asdevice(amplitude)
    ky, kx = _padded_shape(opd.shape, Q)
    kind = OUT_COMPLEX if field else (OUT_INTENSITY if out is None else OUT_ACCUMULATE)
    return _ops.fft2(None, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True,
                     out_kind=kind, weight=weight, out=out, amp=amp, opd=opd, kscale=phase_prefix(wavelength).imag)


def Q_for_sampling(input_diameter, prop_dist, wavelength, output_dx):
    """prysm/propagation/fft.py:88-109."""
    return (wavelength * prop_dist) / input_diameter / output_dx


def pupil_sample_to_psf_sample(pupil_sample, samples, wavelength, efl):
    """prysm/propagation/fft.py:112-132."""
    return (efl * wavelength) / (pupil_sample * samples)


def psf_sample_to_pupil_sample(psf_sample, samples, wavelength, efl):
    """prysm/propagation/fft.py:135-155."""
    return (efl * wavelength) / (psf_sample * samples)


def phase_prefix(wavelength):
    """OPD [nm] -> radians: i*2*pi/wvl/1e3 (prysm/propagation/_kernels.py:40-43)."""
    return 1j * 2 * np.pi / wavelength / 1e3


# ------------------------------------------------------------------------------------------
# array API: angular spectrum (prysm/propagation/angular_spectrum.py)
# ------------------------------------------------------------------------------------------

def angular_spectrum_transfer_function(samples, wvl, dx, z):
    """Full transfer-function array outer(tfy, tfx) (prysm/propagation/angular_spectrum.py:82-114).
    `angular_spectrum` itself never materialises it: it multiplies by the two vectors."""
    if isinstance(samples, int):
        samples = (samples, samples)
    cd = config.complex_dtype
    ty, tx = _ops.angular_spectrum_vectors(tuple(samples), wvl, dx, z, cd, _ops.device())
    ones = torch.ones(tuple(samples), dtype=cd, device=ty.device)
    return _ops.mul_outer(ones, vy=ty, vx=tx)


def angular_spectrum(field, wvl, dx, z, Q=2, tf=None):
    """ifft2(fft2(pad(field)) * tf) in one call; the padded result is returned un-cropped
    (prysm/propagation/angular_spectrum.py:9-42)."""
    f = _field(field)
    if tf is not None:
        return _ops.angular_spectrum(f, tuple(f.shape), tf=_ops.asdevice(tf))
    k = _padded_shape(f.shape, Q)
    ty, tx = _ops.angular_spectrum_vectors(k, wvl, dx, z, f.dtype, f.device)
    return _ops.angular_spectrum(f, k, ty=ty, tx=tx)


def angular_spectrum_adjoint(field, wvl, dx, z, Q=2, tf=None):
    """prysm/propagation/angular_spectrum.py:45-79."""
    g = _field(field)
    k = tuple(g.shape)
    if tf is not None:
        return _ops.angular_spectrum(g, k, tf=_ops.asdevice(tf), conj_tf=True)
    ty, tx = _ops.angular_spectrum_vectors(k, wvl, dx, z, g.dtype, g.device)
    return _ops.angular_spectrum(g, k, ty=ty, tx=tx, conj_tf=True, crop=_shape_before_pad(k, Q))


def fresnel_number(a, L, lambda_):
    """prysm/propagation/angular_spectrum.py:117-138."""
    return a ** 2 / (L * lambda_)


def talbot_distance(a, lambda_):
    """prysm/propagation/angular_spectrum.py:141-164."""
    return lambda_ / (1 - math.sqrt(1 - lambda_ ** 2 / a ** 2))


# ------------------------------------------------------------------------------------------
# array API: fixed-sampling executors (prysm/propagation/dft.py)
# ------------------------------------------------------------------------------------------

def coordinates_for_focus(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl, focal_shift=(0, 0),
                          dtype=None):
    """x, y [mm] and fx, fy [1/mm] as host arrays at config.precision (prysm/propagation/dft.py:12-66).
    `dtype` overrides the precision of the returned vectors."""
    if not isinstance(pupil_samples, Iterable):
        pupil_samples = (pupil_samples, pupil_samples)
    if not isinstance(focal_samples, Iterable):
        focal_samples = (focal_samples, focal_samples)
    pny, pnx = pupil_samples
    fny, fnx = focal_samples
    fsx, fsy = focal_shift
    dt = config.precision if dtype is None else dtype

    def rng(n):
        return np.arange(-(n // 2), -(n // 2) + n, dtype=dt)

    x = rng(pnx) * pupil_dx
    y = rng(pny) * pupil_dx
    inv_lz = 1.0 / (wavelength * efl)
    fx = (rng(fnx) * focal_dx + fsx) * inv_lz
    fy = (rng(fny) * focal_dx + fsy) * inv_lz
    return x, y, fx, fy


def prepare_executor(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl, focal_shift=(0, 0),
                     kind='mdft'):
    """Reusable pupil <-> focal operator with norm = pupil_dx*focal_dx/(wvl*efl) baked in
    (prysm/propagation/dft.py:69-117)."""
    # coordinates stay in fp64 whatever config.precision is: differencing float32 grids (as the
    # reference does at precision=32) costs ~1e-6 in the chirp rate before any transform runs
    x, y, fx, fy = coordinates_for_focus(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl,
                                         focal_shift, dtype=np.float64)
    norm = (pupil_dx * focal_dx) / (wavelength * efl)
    if kind == 'mdft':
        op = MDFT(x, y, fx, fy, sign=-1, norm=norm)
    elif kind == 'czt':
        op = CZT(x, y, fx, fy, sign=-1, norm=norm)
    elif kind == 'fftdft':
        op = FFTDFT(x, y, fx, fy, sign=-1, norm=norm)
    else:
        raise ValueError(f"kind must be 'mdft', 'czt', or 'fftdft', got {kind!r}")
    op.pupil_dx = pupil_dx
    op.focal_dx = focal_dx
    return op


def unit_cell_focal_grid(pupil_dx, pupil_diameter, wavelength, efl, Q=2):
    """prysm/propagation/dft.py:120-152."""
    focal_samples = math.ceil(Q * pupil_diameter / pupil_dx)
    return wavelength * efl / pupil_dx / focal_samples, focal_samples


def focus_dft(wavefunction, executor):
    """prysm/propagation/dft.py:297-313."""
    return executor(wavefunction)


def focus_dft_adjoint(wavefunction, executor):
    """prysm/propagation/dft.py:316-332."""
    return executor.adjoint(wavefunction)


def unfocus_dft(wavefunction, executor):
    """prysm/propagation/dft.py:335-351."""
    return executor.adjoint(wavefunction)


def unfocus_dft_adjoint(wavefunction, executor):
    """prysm/propagation/dft.py:354-370."""
    return executor(wavefunction)


def focus_fixed_sampling(wavefunction, input_dx, prop_dist, wavelength, output_dx, output_samples, shift=(0, 0),
                         method='mdft'):
    """v0.19-v0.21 spelling named by BASELINE.json (docs/source/releases/v0.22.rst:166-205): builds the executor
    and applies it."""
    w = _field(wavefunction)
    ex = prepare_executor(input_dx, tuple(w.shape), output_dx, output_samples, wavelength, prop_dist, shift, method)
    return ex(w)


def unfocus_fixed_sampling(wavefunction, input_dx, prop_dist, wavelength, output_dx, output_samples, shift=(0, 0),
                           method='mdft'):
    """Legacy spelling of unfocus_dft; input_dx is the focal spacing [um], output_dx the pupil spacing [mm]."""
    w = _field(wavefunction)
    ex = prepare_executor(output_dx, output_samples, input_dx, tuple(w.shape), wavelength, prop_dist, shift, method)
    return ex.adjoint(w)


# ------------------------------------------------------------------------------------------
# object API (prysm/propagation/wavefront.py)
# ------------------------------------------------------------------------------------------

class Wavefront:
    """(Complex) representation of a wavefront (prysm/propagation/wavefront.py:35-56)."""

    def __init__(self, cmplx_field, wavelength, dx, space='pupil'):
        self._data = None if cmplx_field is None else _ops.asdevice(cmplx_field)
        self._lazy = None
        self.wavelength = wavelength
        self.dx = dx
        self.space = space

    # `data` materialises a pending from_amp_and_phase on first use
    @property
    def data(self):
        if self._data is None and self._lazy is not None:
            amp, opd = self._lazy
            self._data = _ops.phase_screen(amp, opd, phase_prefix(self.wavelength).imag)
        return self._data

    @data.setter
    def data(self, value):
        self._data = value
        self._lazy = None

    @classmethod
    def from_amp_and_phase(cls, amplitude, phase, wavelength, dx):
        """P = amplitude * exp(i*2*pi/wvl * OPD[nm]) (prysm/propagation/wavefront.py:59-79).
        The complex field is synthesised lazily so that .focus() can fuse it into the FFT."""
        amp = None if amplitude is None else _ops.asdevice(amplitude)
        if phase is None:
            return cls(_ops.ascomplex(amp), wavelength, dx)
        opd = _ops.asdevice(phase)
        if opd.dtype not in (torch.float32, torch.float64):
            opd = opd.to(config.real_dtype)
        out = cls(None, wavelength, dx)
        out._lazy = (amp, opd)
        return out

    @classmethod
    def phase_screen(cls, phase, wavelength, dx):
        """prysm/propagation/wavefront.py:82-96."""
        return cls.from_amp_and_phase(None, phase, wavelength, dx)

    @classmethod
    def thin_lens(cls, f, wavelength, x, y):
        """Quadratic phase exp(-i*2*pi/wvl_mm * r^2/(2f)) (prysm/propagation/wavefront.py:99-144)."""
        x, y = _ops.asdevice(x), _ops.asdevice(y)
        rsq_over_2f = (x * x + y * y) / (2 * f)          # coordinate prep (input generation)
        kscale = -2 * np.pi / (wavelength / 1e3)
        data = _ops.phase_screen(None, rsq_over_2f, kscale)
        dx = float(x[0, 1] - x[0, 0])
        return cls(cmplx_field=data, wavelength=wavelength, dx=dx, space='pupil')

    @property
    def intensity(self):
        """abs(w)^2 as RichData (prysm/propagation/wavefront.py:147-151)."""
        return RichData(_ops.intensity(self.data), self.dx, self.wavelength)

    def copy(self):
        return Wavefront(self.data.clone(), self.wavelength, self.dx, self.space)

    def pad2d(self, Q, value=0, mode='constant', out_shape=None, inplace=True):
        """prysm/propagation/wavefront.py:300-332."""
        padded = pad2d(self.data, Q=Q, value=value, mode=mode, out_shape=out_shape)
        if inplace:
            self.data = padded
            return self
        return Wavefront(padded, self.wavelength, self.dx, self.space)

    def crop(self, out_shape, inplace=True):
        """prysm/propagation/wavefront.py:334-358."""
        cropped = crop_center(self.data, out_shape)
        if inplace:
            self.data = cropped
            return self
        return Wavefront(cropped, self.wavelength, self.dx, self.space)

    def __numerical_operation__(self, other, op, reverse=False):
        """prysm/propagation/wavefront.py:360-379: same physicality checks and exceptions."""
        if isinstance(other, Wavefront):
            criteria = [
                abs(self.dx - other.dx) / self.dx * 100 < 0.1,
                self.data.shape == other.data.shape,
                self.wavelength == other.wavelength,
                self.space == other.space,
            ]
            if not all(criteria):
                raise ValueError('all physicality criteria not met: sample spacing, shape, wavelength, or space different.')
            data = _ops.binary(op, self.data, other.data, reverse)
        elif isinstance(other, torch.Tensor):
            data = _ops.binary(op, self.data, _ops.asdevice(other), reverse)
        elif isinstance(other, numbers.Number):
            data = _ops.binary(op, self.data, other, reverse)
        else:
            raise TypeError(f'unsupported operand type(s) for {op}: \'Wavefront\' and {type(other)}')
        return Wavefront(dx=self.dx, wavelength=self.wavelength, cmplx_field=data, space=self.space)

    def __mul__(self, other):
        return self.__numerical_operation__(other, 'mul')

    def __rmul__(self, other):
        return self.__numerical_operation__(other, 'mul', reverse=True)

    def __truediv__(self, other):
        return self.__numerical_operation__(other, 'truediv')

    def __rtruediv__(self, other):
        return self.__numerical_operation__(other, 'truediv', reverse=True)

    def __add__(self, other):
        return self.__numerical_operation__(other, 'add')

    def __radd__(self, other):
        return self.__numerical_operation__(other, 'add', reverse=True)

    def __sub__(self, other):
        return self.__numerical_operation__(other, 'sub')

    def __rsub__(self, other):
        return self.__numerical_operation__(other, 'sub', reverse=True)

    def free_space(self, dz=np.nan, Q=1, tf=None):
        """Plane-to-plane propagation (prysm/propagation/wavefront.py:413-443)."""
        if np.isnan(dz) and tf is None:
            raise ValueError('dz must be provided if tf is None')
        out = angular_spectrum(self.data, wvl=self.wavelength, dx=self.dx, z=dz, Q=Q, tf=tf)
        return Wavefront(out, self.wavelength, self.dx, self.space)

    def free_space_adjoint(self, dz=np.nan, Q=1, tf=None):
        """prysm/propagation/wavefront.py:445-476."""
        if np.isnan(dz) and tf is None:
            raise ValueError('dz must be provided if tf is None')
        out = angular_spectrum_adjoint(self.data, wvl=self.wavelength, dx=self.dx, z=dz, Q=Q, tf=tf)
        return Wavefront(out, self.wavelength, self.dx, self.space)

    def focus(self, efl, Q=2):
        """Pupil -> psf plane by FFT (prysm/propagation/wavefront.py:478-504)."""
        if self.space != 'pupil':
            raise ValueError('can only propagate from a pupil to psf plane')
        if self._data is None and self._lazy is not None:
            amp, opd = self._lazy
            data = psf_from_amp_and_phase(amp, opd, self.wavelength, Q, field=True)
        else:
            data = focus(self.data, Q=Q)
        dx = pupil_sample_to_psf_sample(self.dx, data.shape[1], self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='psf')

    def focus_adjoint(self, efl, Q=2):
        """prysm/propagation/wavefront.py:506-532."""
        if self.space != 'psf':
            raise ValueError('can only apply adjoint from a psf to pupil plane')
        samples = self.data.shape[1]
        data = focus_adjoint(self.data, Q=Q)
        dx = psf_sample_to_pupil_sample(self.dx, samples, self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='pupil')

    def unfocus(self, efl, Q=2):
        """prysm/propagation/wavefront.py:534-560."""
        if self.space != 'psf':
            raise ValueError('can only propagate from a psf to pupil plane')
        data = unfocus(self.data, Q=Q)
        dx = psf_sample_to_pupil_sample(self.dx, data.shape[1], self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='pupil')

    def unfocus_adjoint(self, efl, Q=2):
        """prysm/propagation/wavefront.py:562-588."""
        if self.space != 'pupil':
            raise ValueError('can only apply adjoint from a pupil to psf plane')
        samples = self.data.shape[1]
        data = unfocus_adjoint(self.data, Q=Q)
        dx = pupil_sample_to_psf_sample(self.dx, samples, self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='psf')

    def _shape(self):
        if self._data is None and self._lazy is not None:
            return tuple(self._lazy[1].shape)
        return tuple(self.data.shape)

    def prepare_executor(self, efl, dx, samples, shift=(0, 0), kind='mdft'):
        """prysm/propagation/wavefront.py:590-641."""
        if isinstance(samples, int):
            samples = (samples, samples)
        if self.space == 'pupil':
            return prepare_executor(pupil_dx=self.dx, pupil_samples=self._shape(), focal_dx=dx,
                                    focal_samples=samples, wavelength=self.wavelength, efl=efl,
                                    focal_shift=shift, kind=kind)
        elif self.space == 'psf':
            return prepare_executor(pupil_dx=dx, pupil_samples=samples, focal_dx=self.dx,
                                    focal_samples=self._shape(), wavelength=self.wavelength, efl=efl,
                                    focal_shift=shift, kind=kind)
        raise ValueError(f"unknown space {self.space!r}")

    def focus_dft(self, executor):
        """prysm/propagation/wavefront.py:679-696."""
        if self.space != 'pupil':
            raise ValueError('can only propagate from a pupil to psf plane')
        data = focus_dft(self.data, executor)
        return Wavefront(dx=executor.focal_dx, cmplx_field=data, wavelength=self.wavelength, space='psf')

    def focus_dft_adjoint(self, executor):
        """prysm/propagation/wavefront.py:698-718."""
        if self.space != 'psf':
            raise ValueError('can only apply adjoint from a psf to pupil plane')
        data = focus_dft_adjoint(self.data, executor)
        return Wavefront(dx=executor.pupil_dx, cmplx_field=data, wavelength=self.wavelength, space='pupil')

    def unfocus_dft(self, executor):
        """prysm/propagation/wavefront.py:720-737."""
        if self.space != 'psf':
            raise ValueError('can only propagate from a psf to pupil plane')
        data = unfocus_dft(self.data, executor)
        return Wavefront(dx=executor.pupil_dx, cmplx_field=data, wavelength=self.wavelength, space='pupil')

    def unfocus_dft_adjoint(self, executor):
        """prysm/propagation/wavefront.py:739-756."""
        if self.space != 'pupil':
            raise ValueError('can only apply adjoint from a pupil to psf plane')
        data = unfocus_dft_adjoint(self.data, executor)
        return Wavefront(dx=executor.focal_dx, cmplx_field=data, wavelength=self.wavelength, space='psf')

    # legacy spellings named by BASELINE.json
    def focus_fixed_sampling(self, efl, dx, samples, shift=(0, 0), method='mdft'):
        return self.focus_dft(self.prepare_executor(efl, dx, samples, shift, method))

    def unfocus_fixed_sampling(self, efl, dx, samples, shift=(0, 0), method='mdft'):
        return self.unfocus_dft(self.prepare_executor(efl, dx, samples, shift, method))
