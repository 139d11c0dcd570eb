"""Image-chain consumers of the path's FFTs on the B200 engine (reference prysm/convolution.py:9-114).

`conv` of a real object with a real PSF runs as ONE forward transform of obj + i*psf, one kernel that separates
the two Hermitian spectra and multiplies them, and ONE inverse transform (the reference does three transforms and a
product); shifts and the 1/(M N) scale are folded into the transforms.  `apply_transfer_functions` multiplies the
object spectrum by each transfer function in one pass per function."""
import inspect

import torch

from . import _ops
from .fttools import forward_ft_unit


def _real_if(obj_is_complex, i):
    return i if obj_is_complex else i.real


def conv(obj, psf):
    """fftshift(ifft2(fft2(ifftshift(obj)) * fft2(ifftshift(psf)))); real for a real object (prysm/convolution.py:9-32)."""
    o, h = _ops.asdevice(obj), _ops.asdevice(psf)
    if tuple(o.shape) != tuple(h.shape):
        raise ValueError(f'obj {tuple(o.shape)} and psf {tuple(h.shape)} must have the same shape')
    ny, nx = o.shape
    if not o.is_complex() and not h.is_complex():
        rd = torch.float64 if torch.float64 in (o.dtype, h.dtype) else torch.float32
        o, h = o.to(rd), h.to(rd)
        s = _ops.balance_scale(o, h)                   # device scalar: the PSF rides at the object's norm
        Z = _ops.fft2(_ops.pack_complex(o, h, s), (ny, nx), dir=-1, shift_in=True)
        OH = _ops.packed_spectrum_product(Z, im_scale=s)
        return _ops.fft2(OH, (ny, nx), dir=+1, scale=1.0 / (ny * nx), shift_out=True).real
    cd = torch.complex128 if torch.float64 in (o.dtype, h.dtype) or torch.complex128 in (o.dtype, h.dtype) else torch.complex64
    O = _ops.fft2(o if not o.is_complex() else o.to(cd), (ny, nx), dir=-1, shift_in=True)
    H = _ops.fft2(h if not h.is_complex() else h.to(cd), (ny, nx), dir=-1, shift_in=True)
    i = _ops.fft2(_ops.binary('mul', O.to(cd), H.to(cd)), (ny, nx), dir=+1, scale=1.0 / (ny * nx), shift_out=True)
    return _real_if(o.is_complex(), i)


def apply_transfer_functions(obj, dx, tfs, fx=None, fy=None, ft=None, fr=None, shift=False):
    """Blur an object by N transfer functions, arrays or callables of any of fx (1, N), fy (M, 1), fr, ft (M, N)
    (prysm/convolution.py:35-114).  Callables receive device tensors."""
    o = _ops.asdevice(obj)
    ny, nx = o.shape
    if any(callable(tf) for tf in tfs):
        if fx is None or fy is None:
            uy, ux = [forward_ft_unit(dx, n, shift=shift) for n in (ny, nx)]
            fx = ux if fx is None else fx
            fy = uy if fy is None else fy
        fx, fy = _ops.asdevice(fx), _ops.asdevice(fy)
        if fx.ndim == 2:                                     # optimize_xy_separable, prysm/coordinates.py:11-46
            fx, fy = fx[0, :], fy[:, 0]
        fx, fy = fx.reshape(1, -1), fy.reshape(-1, 1)
        if fr is None or ft is None:
            cr, ct = _ops.cart_to_polar(fx.expand(fy.shape[0], fx.shape[1]), fy.expand(fy.shape[0], fx.shape[1]))
            fr = cr if fr is None else fr
            ft = ct if ft is None else ft
    O = _ops.fft2(o, (ny, nx), dir=-1, shift_in=True, shift_out=shift)
    for tf in tfs:
        if callable(tf):
            params = inspect.signature(tf).parameters
            kwargs = {k: v for k, v in (('fx', fx), ('fy', fy), ('fr', fr), ('ft', ft)) if k in params}
            if not kwargs:
                raise ValueError(f'{tf} accepts none of fx, fy, fr, ft; a transfer function must accept at least one')
            tf = tf(**kwargs)
        tf = _ops.asdevice(tf)
        if tuple(tf.shape) != (ny, nx):
            tf = tf.expand(ny, nx)
        O = _ops.mask_multiply(O, tf)
    i = _ops.fft2(O, (ny, nx), dir=+1, scale=1.0 / (ny * nx), shift_in=shift, shift_out=True)
    return _real_if(o.is_complex(), i)
