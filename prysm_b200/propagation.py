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
from .coronagraph import (  # noqa: F401  (re-exported like prysm/propagation/__init__.py)
    to_fpm_and_back, to_fpm_and_back_adjoint, to_fpm_and_back_multiresolution,
    to_fpm_and_back_multiresolution_adjoint, babinet, babinet_adjoint, vortex_phase_mask, prepare_measured_fpm,
)


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
    if w.ndim == 3:          # a stack of independent fields (B, ny, nx): one batched call (the reference is 2-D only)
        ky, kx = _padded_shape(w.shape[1:], Q)
        return _ops.fft2_batch(w, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True)
    ky, kx = _padded_shape(w.shape, Q)
    return _ops.fft2(w, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True)


def unfocus(wavefunction, Q):
    """PSF -> pupil plane, the same with ifft2 (prysm/propagation/fft.py:48-65)."""
    w = _field(wavefunction)
    if w.ndim == 3:
        ky, kx = _padded_shape(w.shape[1:], Q)
        return _ops.fft2_batch(w, (ky, kx), dir=+1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True)
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
    if w.ndim == 3:      # a stack of independent fields -> a stack of PSFs in one batched call (no accumulate form)
        if out is not None or weight != 1.0:
            raise ValueError('the batched form writes one |.|^2 plane per field: weight / out are for single fields')
        ky, kx = _padded_shape(w.shape[1:], Q)
        return _ops.fft2_batch(w, (ky, kx), dir=-1, scale=1.0 / math.sqrt(ky * kx), shift_in=True, shift_out=True,
                               out_kind=OUT_INTENSITY)
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


def angular_spectrum(field, wvl, dx, z, Q=2, tf=None, screen=None):
    """ifft2(fft2(pad(field)) * tf) in one call; the padded result is returned un-cropped
    (prysm/propagation/angular_spectrum.py:9-42).  `screen` (an extension): a complex array the field is multiplied
    by inside the first transform pass -- `angular_spectrum(w, ..., screen=s)` == `angular_spectrum(w * s, ...)`."""
    f = _field(field)
    scr = None if screen is None else _field(screen)
    if tf is not None:
        return _ops.angular_spectrum(f, tuple(f.shape), tf=_ops.asdevice(tf), screen=scr)
    k = _padded_shape(f.shape, Q)
    ty, tx = _ops.angular_spectrum_vectors(k, wvl, dx, z, f.dtype, f.device)
    return _ops.angular_spectrum(f, k, ty=ty, tx=tx, screen=scr)


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


class MultiResolutionExecutor:
    """Per-level executors (coarsest first), real partition-of-unity windows and focal coordinate grids [um]
    (prysm/propagation/dft.py:170-209).  Windows and grids are device tensors."""

    __slots__ = ('executors', 'windows', 'xf', 'yf')

    def __init__(self, executors, windows, xf, yf):
        self.executors = executors
        self.windows = windows
        self.xf = xf
        self.yf = yf

    def __len__(self):
        return len(self.executors)


def prepare_multiresolution(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl, num_levels, scaling=4.0,
                            fine_samples=None, window=(0.2, 0.7), kind='mdft'):
    """Stack of executors zooming into the focal origin by `scaling` per level, each shifted by half a focal
    sample, plus the telescoping hand-off windows (prysm/propagation/dft.py:212-294).  Grids and window of a
    level come from one kernel."""
    if fine_samples is None:
        fine_samples = focal_samples
    inner, outer = window
    geo = []
    for k in range(num_levels):
        nf = focal_samples if k == 0 else fine_samples
        if not isinstance(nf, Iterable):
            nf = (nf, nf)
        nfy, nfx = nf
        fdx = focal_dx / scaling ** k
        geo.append((nfy, nfx, fdx, min(nfy, nfx) / 2.0 * fdx))
    executors, windows, xfs, yfs = [], [], [], []
    dev = _ops.device()
    for k, (nfy, nfx, fdx, half) in enumerate(geo):
        shift = fdx / 2.0
        executors.append(prepare_executor(pupil_dx, pupil_samples, fdx, (nfy, nfx), wavelength, efl,
                                          focal_shift=(shift, shift), kind=kind))
        here = None if k == 0 else (inner * half, outer * half)
        nxt = None if k == num_levels - 1 else (inner * geo[k + 1][3], outer * geo[k + 1][3])
        win, xf, yf = _ops.radial_window((nfy, nfx), fdx, shift, here, nxt, config.real_dtype, dev)
        windows.append(win)
        xfs.append(xf)
        yfs.append(yf)
    return MultiResolutionExecutor(executors, windows, xfs, yfs)


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

def _field_data(field):
    """Array of a Wavefront-like, anything else unchanged (prysm/propagation/wavefront.py:28-32)."""
    if isinstance(field, Wavefront):
        return field.data
    return field


class Wavefront:
    """(Complex) representation of a wavefront (prysm/propagation/wavefront.py:35-56)."""

    def __init__(self, cmplx_field, wavelength, dx, space='pupil'):
        self._data = None if cmplx_field is None else _ops.asdevice(cmplx_field)
        self._lazy = None
        self._lazy_mul = None    # (field, screen): a pending elementwise product that free_space() can fuse
        self.wavelength = wavelength
        self.dx = dx
        self.space = space

    # `data` materialises a pending from_amp_and_phase on first use
    @property
    def data(self):
        if self._data is None and self._lazy is not None:
            amp, opd = self._lazy
            self._data = _ops.phase_screen(amp, opd, phase_prefix(self.wavelength).imag)
        elif self._data is None and self._lazy_mul is not None:
            a, b = self._lazy_mul
            self._data = _ops.binary('mul', a, b)
            self._lazy_mul = None
        return self._data

    @data.setter
    def data(self, value):
        self._data = value
        self._lazy = None
        self._lazy_mul = None

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

    @property
    def phase(self):
        """angle(w) (prysm/propagation/wavefront.py:153-156)."""
        return RichData(_ops.component('angle', self.data), self.dx, self.wavelength)

    @property
    def real(self):
        """re(w) (prysm/propagation/wavefront.py:158-161)."""
        return RichData(_ops.component('real', self.data), self.dx, self.wavelength)

    @property
    def imag(self):
        """im(w) (prysm/propagation/wavefront.py:163-166)."""
        return RichData(_ops.component('imag', self.data), self.dx, self.wavelength)

    def copy(self):
        return Wavefront(self.data.clone(), self.wavelength, self.dx, self.space)

    # ---- adjoints of the elementwise constructors (prysm/propagation/wavefront.py:172-298)
    def from_amp_and_phase_adjoint_phase(self, wf_bar):
        """prefix * imag(gbar * conj(g)).  The reference multiplies by the COMPLEX prefix i*2*pi/(1e3*wvl)
        (wavefront.py:187-188), so the result is a purely imaginary array; kept for parity."""
        return _ops.field_adjoint(0, self.data, _field(_field_data(wf_bar)), None, phase_prefix(self.wavelength).imag)

    def from_amp_and_phase_adjoint_amp(self, wf_bar, phase=None):
        """real(gbar * conj(S)), S the unit phasor: rebuilt from `phase` [nm] when given, else P/|P| with zero
        gradient where the amplitude vanishes (prysm/propagation/wavefront.py:190-225)."""
        bar = _field(_field_data(wf_bar))
        k = phase_prefix(self.wavelength).imag
        if phase is not None:
            return _ops.field_adjoint(2, None, bar, _ops.asdevice(phase), k)
        return _ops.field_adjoint(1, self.data, bar, None, k)

    def phase_screen_adjoint_phase(self, wf_bar):
        """prysm/propagation/wavefront.py:227-242."""
        return self.from_amp_and_phase_adjoint_phase(wf_bar)

    @classmethod
    def thin_lens_adjoint(cls, f, wavelength, x, y, wf_bar):
        """Gradient w.r.t. the focal length: pi/(w f^2) * sum(r^2 * imag(Lbar * conj(L))) as one weighted-dot
        reduction (prysm/propagation/wavefront.py:244-280)."""
        L_bar = _field(_field_data(wf_bar))
        L = cls.thin_lens(f, wavelength, x, y).data
        x, y = _ops.asdevice(x), _ops.asdevice(y)
        rsq = x * x + y * y                                   # coordinate prep, as in thin_lens
        w = wavelength / 1e3
        return math.pi / (w * f * f) * _ops.dot(L_bar, L, rsq).imag

    def intensity_adjoint(self, intensity_bar):
        """Gbar = 2 * Ibar * E (prysm/propagation/wavefront.py:282-298)."""
        ibar = intensity_bar.data if isinstance(intensity_bar, (RichData, Wavefront)) else intensity_bar
        ibar = _ops.asdevice(ibar)
        return Wavefront(_ops.mask_multiply(self.data, ibar, scale=2.0), self.wavelength, self.dx, self.space)

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
            if op == 'mul':   # kept pending: `(wf * screen).free_space(dz)` multiplies inside the first transform pass
                out = Wavefront(None, self.wavelength, self.dx, self.space)
                out._lazy_mul = (self.data, _ops.ascomplex(other.data).to(self.data.dtype))
                return out
            data = _ops.binary(op, self.data, other.data, reverse)
        elif isinstance(other, (torch.Tensor, np.ndarray)):      # host arrays are uploaded like everywhere else
            od = _ops.asdevice(other)
            if op == 'mul' and tuple(od.shape) == tuple(self.data.shape):
                out = Wavefront(None, self.wavelength, self.dx, self.space)
                out._lazy_mul = (self.data, _ops.ascomplex(od).to(self.data.dtype))
                return out
            data = _ops.binary(op, self.data, od, reverse)
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
        if self._data is None and self._lazy_mul is not None:     # pending wf * screen: fused into the first pass
            a, b = self._lazy_mul
            out = angular_spectrum(a, wvl=self.wavelength, dx=self.dx, z=dz, Q=Q, tf=tf, screen=b)
        else:
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
        dx = pupil_sample_to_psf_sample(self.dx, data.shape[-1], self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='psf')

    def focus_adjoint(self, efl, Q=2):
        """prysm/propagation/wavefront.py:506-532."""
        if self.space != 'psf':
            raise ValueError('can only apply adjoint from a psf to pupil plane')
        samples = self.data.shape[-1]
        data = focus_adjoint(self.data, Q=Q)
        dx = psf_sample_to_pupil_sample(self.dx, samples, self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='pupil')

    def unfocus(self, efl, Q=2):
        """prysm/propagation/wavefront.py:534-560."""
        if self.space != 'psf':
            raise ValueError('can only propagate from a psf to pupil plane')
        data = unfocus(self.data, Q=Q)
        dx = psf_sample_to_pupil_sample(self.dx, data.shape[-1], self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='pupil')

    def unfocus_adjoint(self, efl, Q=2):
        """prysm/propagation/wavefront.py:562-588."""
        if self.space != 'pupil':
            raise ValueError('can only apply adjoint from a pupil to psf plane')
        samples = self.data.shape[-1]
        data = unfocus_adjoint(self.data, Q=Q)
        dx = pupil_sample_to_psf_sample(self.dx, samples, self.wavelength, efl)
        return Wavefront(data, self.wavelength, dx, space='psf')

    def _shape(self):
        if self._data is None and self._lazy is not None:
            return tuple(self._lazy[1].shape)
        if self._data is None and self._lazy_mul is not None:
            return tuple(self._lazy_mul[0].shape)
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

    # ---- Lyot-coronagraph compositions (prysm/propagation/wavefront.py:643-677, 758-1038)
    def prepare_multiresolution(self, efl, focal_dx, focal_samples, num_levels, scaling=4.0, fine_samples=None,
                                window=(0.2, 0.7), kind='mdft'):
        if self.space != 'pupil':
            raise ValueError('multiresolution propagation begins at a pupil plane')
        return prepare_multiresolution(pupil_dx=self.dx, pupil_samples=self._shape(), focal_dx=focal_dx,
                                       focal_samples=focal_samples, wavelength=self.wavelength, efl=efl,
                                       num_levels=num_levels, scaling=scaling, fine_samples=fine_samples,
                                       window=window, kind=kind)

    def _pupil(self, data):
        return Wavefront(data, self.wavelength, self.dx, self.space)

    def _psf(self, data, executor):
        return Wavefront(data, self.wavelength, executor.focal_dx, 'psf')

    def to_fpm_and_back(self, fpm, executor, return_more=False):
        pak = to_fpm_and_back(self.data, fpm=_field_data(fpm), executor=executor, return_more=return_more)
        if return_more:
            nxt, at_fpm, after_fpm = pak
            return self._pupil(nxt), self._psf(at_fpm, executor), self._psf(after_fpm, executor)
        return self._pupil(pak)

    def to_fpm_and_back_adjoint(self, fpm, executor, return_more=False, return_fpm_grad=False, field_at_fpm=None):
        pak = to_fpm_and_back_adjoint(self.data, fpm=_field_data(fpm), executor=executor, return_more=return_more,
                                      return_fpm_grad=return_fpm_grad, field_at_fpm=_field_data(field_at_fpm))
        if not (return_more or return_fpm_grad):
            return self._pupil(pak)
        return (self._pupil(pak[0]),) + tuple(self._psf(p, executor) for p in pak[1:])

    def to_fpm_and_back_multiresolution(self, fpm, executor, return_more=False):
        if self.space != 'pupil':
            raise ValueError('can only propagate from a pupil to psf plane')
        pak = to_fpm_and_back_multiresolution(self.data, fpm, executor, return_more=return_more)
        if not return_more:
            return self._pupil(pak)
        out, at_fpm, after_fpm = pak
        return (self._pupil(out), [self._psf(f, ex) for f, ex in zip(at_fpm, executor.executors)],
                [self._psf(f, ex) for f, ex in zip(after_fpm, executor.executors)])

    def to_fpm_and_back_multiresolution_adjoint(self, fpm, executor, return_more=False, return_fpm_grad=False,
                                                field_at_fpm=None):
        if field_at_fpm is not None:
            field_at_fpm = [_field_data(f) for f in field_at_fpm]
        pak = to_fpm_and_back_multiresolution_adjoint(self.data, fpm, executor, return_more=return_more,
                                                      return_fpm_grad=return_fpm_grad, field_at_fpm=field_at_fpm)
        if not (return_more or return_fpm_grad):
            return self._pupil(pak)
        return (self._pupil(pak[0]),) + tuple([self._psf(f, ex) for f, ex in zip(fields, executor.executors)]
                                               for fields in pak[1:])

    def babinet(self, lyot, fpm, executor, return_more=False):
        pak = babinet(self.data, lyot=_field_data(lyot), fpm=_field_data(fpm), executor=executor,
                      return_more=return_more)
        if return_more:
            after_lyot, at_fpm, after_fpm, at_lyot = pak
            return (self._pupil(after_lyot), self._psf(at_fpm, executor), self._psf(after_fpm, executor),
                    self._pupil(at_lyot))
        return self._pupil(pak)

    def babinet_adjoint(self, lyot, fpm, executor, field_at_fpm=None, field_at_lyot=None, return_fpm_grad=False,
                        return_lyot_grad=False):
        pak = babinet_adjoint(self.data, lyot=_field_data(lyot), fpm=_field_data(fpm), executor=executor,
                              field_at_fpm=_field_data(field_at_fpm), field_at_lyot=_field_data(field_at_lyot),
                              return_fpm_grad=return_fpm_grad, return_lyot_grad=return_lyot_grad)
        if not (return_fpm_grad or return_lyot_grad):
            return self._pupil(pak)
        pak = list(pak)
        out = [self._pupil(pak.pop(0))]
        if return_fpm_grad:
            out.append(self._psf(pak.pop(0), executor))
        if return_lyot_grad:
            out.append(self._pupil(pak.pop(0)))
        return tuple(out)

    # legacy spellings named by BASELINE.json
    def focus_fixed_sampling(self, efl, dx, samples, shift=(0, 0), method='mdft'):
        return self.focus_dft(self.prepare_executor(efl, dx, samples, shift, method))

    def unfocus_fixed_sampling(self, efl, dx, samples, shift=(0, 0), method='mdft'):
        return self.unfocus_dft(self.prepare_executor(efl, dx, samples, shift, method))
