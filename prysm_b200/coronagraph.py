"""Lyot-family coronagraph compositions on the B200 engine (reference prysm/propagation/coronagraph.py).

Same names, arguments, return tuples and exceptions as the reference.  Each composition is the executor's
forward / adjoint transform (tensor-core MDFT, CZT or FFTDFT -- prysm_b200.fttools) around ONE masked-multiply
kernel per elementwise step: the Babinet complement 1 - fpm, the partition-of-unity window, the conjugation of
the adjoint, the real-part of a real-mask gradient, the subtraction at the Lyot plane and the level sum are
folded into that kernel instead of being separate full-array passes.
"""
import numbers

import torch

from . import _ops


def _field(w):
    return _ops.ascomplex(_ops.asdevice(w))


def _adjoint_multiply(grad, factor, real=False, window=None):
    """Adjoint w.r.t. x of y = x * factor (* window): grad * conj(factor) (prysm/propagation/_kernels.py:30-38)."""
    return _ops.mask_multiply(_field(grad), _ops.asdevice(factor), conj=True, real_out=real, w=window)


def to_fpm_and_back(wavefunction, fpm, executor, return_more=False):
    """focus_dft -> multiply by fpm -> unfocus_dft with one executor (prysm/propagation/coronagraph.py:12-46)."""
    field_at_fpm = executor(_field(wavefunction))
    field_after_fpm = _ops.mask_multiply(field_at_fpm, _ops.asdevice(fpm))
    field_at_next_pupil = executor.adjoint(field_after_fpm)
    if return_more:
        return field_at_next_pupil, field_at_fpm, field_after_fpm
    return field_at_next_pupil


def to_fpm_and_back_adjoint(wavefunction, fpm, executor, return_more=False, return_fpm_grad=False, field_at_fpm=None):
    """Adjoint of to_fpm_and_back (prysm/propagation/coronagraph.py:49-99)."""
    if return_fpm_grad and field_at_fpm is None:
        raise ValueError('return_fpm_grad=True requires field_at_fpm from the forward propagation')
    fpm = _ops.asdevice(fpm)
    Ebbar = executor(_field(wavefunction))                    # unfocus_dft_adjoint
    intermediate = _adjoint_multiply(Ebbar, fpm)
    Eabar = executor.adjoint(intermediate)                    # focus_dft_adjoint
    if return_fpm_grad:
        fpm_bar = _adjoint_multiply(Ebbar, _field(field_at_fpm), real=not fpm.is_complex())
    if return_more:
        if return_fpm_grad:
            return Eabar, Ebbar, intermediate, fpm_bar
        return Eabar, Ebbar, intermediate
    if return_fpm_grad:
        return Eabar, fpm_bar
    return Eabar


class _VortexMask:
    """fpm(xf, yf) = exp(i*charge*atan2(yf, xf)) evaluated by one kernel on device coordinate grids."""

    def __init__(self, charge):
        self.charge = int(charge)

    def __call__(self, xf, yf):
        xf = _ops.asdevice(xf)
        yf = _ops.asdevice(yf)
        if xf.dtype not in (torch.float32, torch.float64):
            xf = xf.to(torch.float64)
        return _ops.vortex_phase(self.charge, xf, yf)


def vortex_phase_mask(charge):
    """Focal-plane-mask callable of a charge-`charge` optical vortex (prysm/propagation/coronagraph.py:102-132)."""
    if not isinstance(charge, numbers.Integral):
        raise TypeError(f'charge must be an integer, got {charge!r}; non-integer charge has a branch cut at theta=pi')
    return _VortexMask(charge)


def prepare_measured_fpm(measurement, dx, center=(0, 0), charge=None, fill=None, order=1):
    """Wrap a measured complex focal-plane-mask map as an fpm(xf, yf) callable for the multi-resolution executor
    (prysm/propagation/coronagraph.py:135-209): each level resamples the map at its own focal grid in one kernel --
    bilinear (order=1, the reference's default; higher spline orders are not on the device), edge-replicating, with
    `fill` (scalar or callable; default an ideal charge-`charge` vortex, else 1) outside the measured extent."""
    if order != 1:
        raise NotImplementedError('only order=1 (bilinear, the reference default) resampling runs on the device')
    meas = _field(measurement)
    cx, cy = center
    if fill is None:
        fill = vortex_phase_mask(charge) if charge is not None else 1.0

    def fpm(xf, yf):
        xf = _ops.asdevice(xf)
        yf = _ops.asdevice(yf)
        if xf.dtype not in (torch.float32, torch.float64):
            xf = xf.to(torch.float64)
        fillv = _ops.asdevice(fill(xf, yf)) if callable(fill) else fill
        return _ops.resample_bilinear(meas, xf, yf, cx, cy, dx, fillv)

    return fpm


def to_fpm_and_back_multiresolution(wavefunction, fpm, executor, return_more=False):
    """Sum over levels of unfocus_dft(focus_dft(w) * fpm(xf, yf) * window) (prysm/propagation/coronagraph.py:212-251).
    The mask, the window and the product are one pass; the level sum is a fused accumulate."""
    w = _field(wavefunction)
    out = None
    fields_at_fpm = []
    fields_after_fpm = []
    for ex, win, xf, yf in zip(executor.executors, executor.windows, executor.xf, executor.yf):
        field_at_fpm = ex(w)
        field_after_fpm = _ops.mask_multiply(field_at_fpm, _ops.asdevice(fpm(xf, yf)), w=win)
        contribution = ex.adjoint(field_after_fpm)
        out = contribution if out is None else _ops.binary('add', out, contribution)
        if return_more:
            fields_at_fpm.append(field_at_fpm)
            fields_after_fpm.append(field_after_fpm)
    if return_more:
        return out, fields_at_fpm, fields_after_fpm
    return out


def to_fpm_and_back_multiresolution_adjoint(wavefunction, fpm, executor, return_more=False, return_fpm_grad=False,
                                            field_at_fpm=None):
    """Adjoint of to_fpm_and_back_multiresolution (prysm/propagation/coronagraph.py:254-298)."""
    if return_fpm_grad and field_at_fpm is None:
        raise ValueError('return_fpm_grad=True requires field_at_fpm from the forward propagation')
    g = _field(wavefunction)
    out = None
    Ebbars = []
    intermediates = []
    fpm_bars = []
    levels = zip(executor.executors, executor.windows, executor.xf, executor.yf)
    for k, (ex, win, xf, yf) in enumerate(levels):
        m = _ops.asdevice(fpm(xf, yf))
        Ebbar = ex(g)
        intermediate = _adjoint_multiply(Ebbar, m, window=win)
        contribution = ex.adjoint(intermediate)
        out = contribution if out is None else _ops.binary('add', out, contribution)
        if return_more:
            Ebbars.append(Ebbar)
            intermediates.append(intermediate)
        if return_fpm_grad:
            fpm_bars.append(_adjoint_multiply(Ebbar, _field(field_at_fpm[k]), real=not m.is_complex(), window=win))
    if return_more:
        if return_fpm_grad:
            return out, Ebbars, intermediates, fpm_bars
        return out, Ebbars, intermediates
    if return_fpm_grad:
        return out, fpm_bars
    return out


def babinet(wavefunction, lyot, fpm, executor, return_more=False):
    """lyot * (w - to_fpm_and_back(w, 1 - fpm)) (prysm/propagation/coronagraph.py:301-353).  The complement is
    formed inside the mask kernel and the Lyot-plane subtract + stop multiply are one pass."""
    w = _field(wavefunction)
    field_at_fpm = executor(w)
    field_after_fpm = _ops.mask_multiply(field_at_fpm, _ops.asdevice(fpm), one_minus=True)
    field = executor.adjoint(field_after_fpm)
    if return_more or lyot is None:
        field_at_lyot = _ops.binary('sub', w, field)
        field_after_lyot = field_at_lyot if lyot is None else _ops.mask_multiply(field_at_lyot, _ops.asdevice(lyot))
    else:
        field_after_lyot = _ops.mask_multiply(w, _ops.asdevice(lyot), b=field)
    if return_more:
        return field_after_lyot, field_at_fpm, field_after_fpm, field_at_lyot
    return field_after_lyot


def babinet_adjoint(wavefunction, lyot, fpm, executor, field_at_fpm=None, field_at_lyot=None, return_fpm_grad=False,
                    return_lyot_grad=False):
    """Adjoint of babinet (prysm/propagation/coronagraph.py:356-431)."""
    if return_lyot_grad and field_at_lyot is None:
        raise ValueError('return_lyot_grad=True requires field_at_lyot from the forward propagation')
    if return_fpm_grad and field_at_fpm is None:
        raise ValueError('return_fpm_grad=True requires field_at_fpm from the forward propagation')
    dbar = _field(wavefunction)
    fpm = _ops.asdevice(fpm)
    lyot = None if lyot is None else _ops.asdevice(lyot)
    lyot_is_complex = True if lyot is None else lyot.is_complex()
    cbar = dbar if lyot is None else _adjoint_multiply(dbar, lyot)
    Ebbar = executor(cbar)
    intermediate = _ops.mask_multiply(Ebbar, fpm, conj=True, one_minus=True)      # * conj(1 - fpm)
    abar = _ops.binary('sub', cbar, executor.adjoint(intermediate))
    if not (return_fpm_grad or return_lyot_grad):
        return abar
    out = [abar]
    if return_fpm_grad:
        # d/dfpm of lyot*(w - c(1 - fpm)): the two sign flips cancel, so this is c's gradient w.r.t. its mask
        out.append(_adjoint_multiply(Ebbar, _field(field_at_fpm), real=not fpm.is_complex()))
    if return_lyot_grad:
        out.append(_adjoint_multiply(dbar, _field(field_at_lyot), real=not lyot_is_complex))
    return tuple(out)
