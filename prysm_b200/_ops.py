"""Thin tensor-level wrappers over the C ABI: each function takes CUDA torch tensors (or host
arrays, which are uploaded), allocates the output with torch's caching allocator and makes
exactly one library call on torch's current stream.  PyTorch is plumbing here (device memory,
streams); every kernel is in libprysm_b200.so.
"""
import math
import ctypes as C

import numpy as np
import torch

from . import _capi as capi
from ._capi import lib

_REAL_OF = {torch.complex64: torch.float32, torch.complex128: torch.float64}
_CPLX_OF = {torch.float32: torch.complex64, torch.float64: torch.complex128}
_CODE = {torch.complex64: capi.PB_C64, torch.complex128: capi.PB_C128,
         torch.float32: capi.PB_C64, torch.float64: capi.PB_C128}

_default_device = None


def set_device(device=None):
    """Select the CUDA device new arrays are placed on (default: torch's current device)."""
    global _default_device
    _default_device = None if device is None else torch.device(device)


def device():
    if _default_device is not None:
        return _default_device
    if not torch.cuda.is_available():
        raise capi.B200Error('no CUDA device is visible: prysm_b200 has no CPU fallback')
    return torch.device('cuda', torch.cuda.current_device())


def torch_dtype(np_or_torch_dtype):
    if isinstance(np_or_torch_dtype, torch.dtype):
        return np_or_torch_dtype
    return {np.dtype('float32'): torch.float32, np.dtype('float64'): torch.float64,
            np.dtype('complex64'): torch.complex64, np.dtype('complex128'): torch.complex128,
            np.dtype('bool'): torch.bool, np.dtype('uint8'): torch.uint8,
            np.dtype('int64'): torch.int64, np.dtype('int32'): torch.int32}[np.dtype(np_or_torch_dtype)]


def asdevice(a, dtype=None):
    """Host array / scalar sequence / tensor -> contiguous CUDA tensor (no copy if already one)."""
    if isinstance(a, torch.Tensor):
        t = a if a.is_cuda else a.to(device(), non_blocking=True)
    else:
        t = torch.as_tensor(np.ascontiguousarray(a)).to(device(), non_blocking=True)
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    return t.contiguous()


def asnumpy(t):
    return t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)


def ascomplex(t):
    """Promote a real / bool tensor to the complex dtype of matching precision."""
    if t.is_complex():
        return t
    if t.dtype == torch.float64:
        return t.to(torch.complex128)
    return t.to(torch.complex64)


def _ctx(t):
    stream = torch.cuda.current_stream(t.device).cuda_stream
    return capi.handle_for(t.device.index, stream), C.c_void_p(stream)


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _check_out(out, shape, dtype, device, name='out'):
    """A caller-supplied output goes to the kernel as a raw pointer + row pitch: refuse anything the kernel would
    misread (wrong dtype / shape / device, rows that are not unit-stride) instead of writing out of bounds."""
    if not isinstance(out, torch.Tensor):
        raise TypeError(f'`{name}` must be a torch.Tensor, got {type(out).__name__}')
    if out.dtype != dtype:
        raise ValueError(f'`{name}` has dtype {out.dtype}, the operation writes {dtype}')
    if tuple(out.shape) != tuple(shape):
        raise ValueError(f'`{name}` has shape {tuple(out.shape)}, the operation writes {tuple(shape)}')
    if out.device != device:
        raise ValueError(f'`{name}` is on {out.device}, the input is on {device}')
    if out.ndim and out.stride(-1) != 1:
        raise ValueError(f'`{name}` must have unit stride along its last axis')
    return out


def _darr(v):
    v = np.ascontiguousarray(v, dtype=np.float64)
    return v, v.ctypes.data_as(C.POINTER(C.c_double))


# ------------------------------------------------------------------------------------------
def fft2(field, k=None, *, dir=-1, scale=1.0, shift_in=False, shift_out=False, crop=None,
         out_kind=capi.OUT_COMPLEX, weight=1.0, out=None, amp=None, opd=None, kscale=0.0):
    """pb_fft2.  Either `field` (complex or real 2-D tensor) or (amp, opd, kscale)."""
    if opd is not None:
        src = opd
        in_kind = capi.IN_AMP_OPD
        if amp is None:
            amp_kind, amp_t = capi.AMP_NONE, None
        elif amp.dtype in (torch.bool, torch.uint8):
            amp_kind, amp_t = capi.AMP_U8, amp.contiguous()
        else:
            amp_kind, amp_t = capi.AMP_REAL, amp.to(opd.dtype).contiguous()
        cdtype = _CPLX_OF[opd.dtype]
    else:
        src = field
        amp_kind, amp_t = capi.AMP_NONE, None
        if field.is_complex():
            in_kind, cdtype = capi.IN_COMPLEX, field.dtype
        else:
            if field.dtype not in _CPLX_OF:
                src = field.to(torch.float32)
            in_kind, cdtype = capi.IN_REAL, _CPLX_OF[src.dtype]
    src = src.contiguous()
    ny, nx = src.shape
    ky, kx = (ny, nx) if k is None else k
    oy, ox = (ky, kx) if crop is None else crop
    odt = cdtype if out_kind == capi.OUT_COMPLEX else _REAL_OF[cdtype]
    if out is None:
        if out_kind == capi.OUT_ACCUMULATE:
            raise ValueError('accumulate needs an existing `out` array')
        out = torch.empty((oy, ox), dtype=odt, device=src.device)
    else:
        _check_out(out, (oy, ox), odt, src.device)
    h, st = _ctx(src)
    h.check(lib.pb_fft2(h.ptr, _CODE[cdtype], _p(src), in_kind, _p(amp_t), amp_kind, float(kscale),
                        ny, nx, nx, ky, kx, int(dir), float(scale), int(shift_in), int(shift_out),
                        _p(out), out_kind, float(weight), oy, ox, out.stride(0), st))
    return out


def fft2_batch(fields, k=None, *, dir=-1, scale=1.0, shift_in=False, shift_out=False, out_kind=capi.OUT_COMPLEX, out=None,
               amp=None, opd=None, kscale=0.0):
    """pb_fft2_batch over a (B, ny, nx) stack -> (B, ky, kx): either `fields` (complex or real) or `opd` (B, ny, nx)
    with an optional amplitude that is (ny, nx) (shared) or (B, ny, nx)."""
    amp_kind, amp_t, amp_bs = capi.AMP_NONE, None, 0
    if opd is not None:
        src = opd if opd.dtype in _CPLX_OF else opd.to(torch.float32)
        in_kind, cdtype = capi.IN_AMP_OPD, _CPLX_OF[src.dtype]
        if amp is not None:
            if amp.dtype in (torch.bool, torch.uint8):
                amp_kind, amp_t = capi.AMP_U8, amp.contiguous()
            else:
                amp_kind, amp_t = capi.AMP_REAL, amp.to(src.dtype).contiguous()
            amp_bs = amp_t.stride(0) if amp_t.ndim == 3 else 0
    else:
        src = fields
        if src.is_complex():
            in_kind, cdtype = capi.IN_COMPLEX, src.dtype
        else:
            if src.dtype not in _CPLX_OF:
                src = src.to(torch.float32)
            in_kind, cdtype = capi.IN_REAL, _CPLX_OF[src.dtype]
    src = src.contiguous()
    nb, ny, nx = src.shape
    ky, kx = (ny, nx) if k is None else k
    odt = cdtype if out_kind == capi.OUT_COMPLEX else _REAL_OF[cdtype]
    if out is None:
        out = torch.empty((nb, ky, kx), dtype=odt, device=src.device)
    else:
        _check_out(out, (nb, ky, kx), odt, src.device)
        if out.stride(1) < kx or (nb > 1 and out.stride(0) < (ky - 1) * out.stride(1) + kx):
            raise ValueError('`out` rows / fields overlap')
    h, st = _ctx(src)
    h.check(lib.pb_fft2_batch(h.ptr, _CODE[cdtype], _p(src), in_kind, _p(amp_t), amp_kind, float(kscale), nb,
                              src.stride(0), amp_bs, ny, nx, src.stride(1), ky, kx, int(dir), float(scale), int(shift_in),
                              int(shift_out), _p(out), out_kind, 1.0, ky, kx, out.stride(1), out.stride(0), st))
    return out


def fft1(a, n=None, axis=-1, dir=-1, scale=1.0):
    """pb_fft1: numpy fft(a, n, axis) on a 2-D complex tensor."""
    a = ascomplex(a).contiguous()
    ny, nx = a.shape
    axis = axis % 2
    if n is None:
        n = a.shape[axis]
    oshape = (n, nx) if axis == 0 else (ny, n)
    out = torch.empty(oshape, dtype=a.dtype, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_fft1(h.ptr, _CODE[a.dtype], _p(a), ny, nx, nx, axis, int(n), int(dir), float(scale),
                        _p(out), oshape[1], st))
    return out


def axis_dft(a, n, axis, dir=-1, scale=1.0, pre_e=None, pre_e_conj=False, pre_b=None, pre_b_conj=False,
             post_e=None, post_e_conj=False, post_b=None, post_b_conj=False, out_off=0, n_out=None):
    """pb_axis_dft: one DFT pass with fused multipliers and an output window."""
    a = ascomplex(a).contiguous()
    ny, nx = a.shape
    axis = axis % 2
    if n_out is None:
        n_out = n - out_off
    oshape = (n_out, nx) if axis == 0 else (ny, n_out)
    out = torch.empty(oshape, dtype=a.dtype, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_axis_dft(h.ptr, _CODE[a.dtype], _p(a), ny, nx, nx, axis, int(n), int(dir), float(scale),
                            _p(pre_e), int(pre_e_conj), _p(pre_b), int(pre_b_conj),
                            _p(post_e), int(post_e_conj), _p(post_b), int(post_b_conj),
                            int(out_off), int(n_out), _p(out), oshape[1], st))
    return out


def czt_plan(N, M, K, shift, alpha, sign, xc, f0, df, cdtype, dev):
    """pb_czt_plan: (b, post, H, Hadj) for one axis, built on the device."""
    b = torch.empty(N, dtype=cdtype, device=dev)
    post = torch.empty(M, dtype=cdtype, device=dev)
    H = torch.empty(K, dtype=cdtype, device=dev)
    Hadj = torch.empty(K, dtype=cdtype, device=dev)
    h, st = _ctx(b)
    h.check(lib.pb_czt_plan(h.ptr, _CODE[cdtype], int(N), int(M), int(K), float(shift), float(alpha), int(sign), float(xc),
                            float(f0), float(df), _p(b), _p(post), _p(H), _p(Hadj), st))
    return b, post, H, Hadj


def polychromatic_czt(amp, opd, m, K, units, plane):
    """pb_polychromatic_czt: plane += sum_i weight_i |CZT_i(amp exp(i kscale_i opd))|^2 for the rows of `units`
    (host float64 array (n_units, 8): kscale, shift, alpha, xc, f0, df, norm, weight) -- one library call for the loop."""
    import numpy as np
    opd = opd.contiguous()
    n = opd.shape[0]
    if opd.ndim != 2 or opd.shape[1] != n:
        raise ValueError('a square OPD array is required')
    rdt = opd.dtype
    if rdt not in _CPLX_OF:
        raise ValueError('opd must be float32 or float64')
    if amp is None:
        amp_kind, amp_t = capi.AMP_NONE, None
    elif amp.dtype in (torch.bool, torch.uint8):
        amp_kind, amp_t = capi.AMP_U8, amp.contiguous()
    else:
        amp_kind, amp_t = capi.AMP_REAL, amp.to(rdt).contiguous()
    if amp_t is not None and tuple(amp_t.shape) != (n, n):
        raise ValueError('amplitude and OPD shapes differ')
    _check_out(plane, (m, m), rdt, opd.device, name='plane')
    if not plane.is_contiguous():
        raise ValueError('`plane` must be contiguous')
    units = np.ascontiguousarray(units, dtype=np.float64)
    if units.ndim != 2 or units.shape[1] != 8:
        raise ValueError('units must have shape (n_units, 8)')
    code = _CODE[_CPLX_OF[rdt]]
    nbytes = int(lib.pb_polychromatic_czt_work_bytes(code, n, int(m), int(K)))
    work = torch.empty(nbytes + 256, dtype=torch.uint8, device=opd.device)
    off = (-work.data_ptr()) % 256
    h, st = _ctx(opd)
    import ctypes
    h.check(lib.pb_polychromatic_czt(h.ptr, code, _p(amp_t), amp_kind, _p(opd), n, int(m), int(K), int(units.shape[0]),
                                     units.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), work.data_ptr() + off, _p(plane), st))
    return plane


def czt_axis(a, K, axis, pre_e, H, post_e, out_off, n_out, scale=1.0, pre_conj=False, post_conj=False,
             out_kind=capi.OUT_COMPLEX, weight=1.0, out=None):
    """pb_czt_axis: (a*pre_e) -> FFT_K -> *H -> IFFT_K -> [out_off:out_off+n_out] -> *post_e*scale along `axis`.
    `a` may be a row-pitched view (a column slice of a larger array): only unit column stride is required."""
    a = ascomplex(a)
    if a.stride(1) != 1 or a.stride(0) < a.shape[1]:
        a = a.contiguous()
    ny, nx = a.shape
    axis = axis % 2
    oshape = (n_out, nx) if axis == 0 else (ny, n_out)
    h, st = _ctx(a)
    if out_kind != capi.OUT_COMPLEX:     # |.|^2 (written or accumulated with `weight`) formed in the pass's store
        rd = _REAL_OF[a.dtype]
        if out is None:
            if out_kind == capi.OUT_ACCUMULATE:
                raise ValueError('accumulate needs an existing `out` array')
            out = torch.empty(oshape, dtype=rd, device=a.device)
        else:
            _check_out(out, oshape, rd, a.device)
        h.check(lib.pb_czt_axis_intensity(h.ptr, _CODE[a.dtype], _p(a), ny, nx, a.stride(0), axis, int(K), _p(pre_e),
                                          int(pre_conj), _p(H), _p(post_e), int(post_conj), int(out_off), int(n_out),
                                          float(scale), int(out_kind), float(weight), _p(out), out.stride(0), st))
        return out
    out = torch.empty(oshape, dtype=a.dtype, device=a.device)
    h.check(lib.pb_czt_axis(h.ptr, _CODE[a.dtype], _p(a), ny, nx, a.stride(0), axis, int(K), _p(pre_e), int(pre_conj), _p(H),
                            _p(post_e), int(post_conj), int(out_off), int(n_out), float(scale), _p(out), oshape[1], st))
    return out


def angular_spectrum(field, k, ty=None, tx=None, tf=None, conj_tf=False, crop=None, screen=None, conj_screen=False):
    """pb_angular_spectrum; with `screen` (complex, same shape as the field) pb_angular_spectrum_screen: the field is
    multiplied by the screen inside the first transform pass."""
    field = ascomplex(field).contiguous()
    ny, nx = field.shape
    ky, kx = k
    oy, ox = (ky, kx) if crop is None else crop
    out = torch.empty((oy, ox), dtype=field.dtype, device=field.device)
    if tf is not None:
        tf = tf.to(field.dtype).contiguous()
    h, st = _ctx(field)
    if screen is not None:
        if tuple(screen.shape) != (ny, nx):
            raise ValueError(f'shape mismatch {(ny, nx)} vs {tuple(screen.shape)}')
        screen = ascomplex(screen).to(field.dtype).contiguous()
        h.check(lib.pb_angular_spectrum_screen(h.ptr, _CODE[field.dtype], _p(field), _p(screen), int(conj_screen), ny, nx, ky,
                                               kx, _p(ty), _p(tx), _p(tf), int(conj_tf), _p(out), oy, ox, st))
        return out
    h.check(lib.pb_angular_spectrum(h.ptr, _CODE[field.dtype], _p(field), ny, nx, ky, kx, _p(ty), _p(tx), _p(tf),
                                    int(conj_tf), _p(out), oy, ox, st))
    return out


_AS_VECTORS = {}     # (shape, wvl, dx, z, dtype, device, stream) -> (ty, tx); most recently used last


def angular_spectrum_vectors(shape, wvl, dx, z, cdtype, dev):
    """The two separable factors of the free-space transfer function (pb_angular_spectrum_vectors).  A plane-to-plane chain
    asks for the same pair at every step (same grid, wavelength and dz): the last few pairs are kept per stream (they are
    read-only, 8 or 16 bytes per sample of one axis), which takes two small launches out of every step of the chain.
    Skipped while a CUDA graph is being captured (a captured graph must own the launches that fill what it reads)."""
    ky, kx = shape
    dev = torch.device(dev)
    capturing = torch.cuda.is_current_stream_capturing() if dev.type == 'cuda' else False
    key = None
    if not capturing and dev.type == 'cuda':
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        key = (int(ky), int(kx), float(wvl), float(dx), float(z), cdtype, idx, torch.cuda.current_stream(dev).cuda_stream)
        hit = _AS_VECTORS.pop(key, None)
        if hit is not None:
            _AS_VECTORS[key] = hit
            return hit
    ty = torch.empty(ky, dtype=cdtype, device=dev)
    tx = torch.empty(kx, dtype=cdtype, device=dev)
    h, st = _ctx(ty)
    h.check(lib.pb_angular_spectrum_vectors(h.ptr, _CODE[cdtype], ky, kx, float(wvl), float(dx), float(z),
                                            _p(ty), _p(tx), st))
    if key is not None:
        _AS_VECTORS[key] = (ty, tx)
        while len(_AS_VECTORS) > 8:
            _AS_VECTORS.pop(next(iter(_AS_VECTORS)))
    return ty, tx


def mdft_basis(f, x, sign, cdtype, dev):
    """E[j, l] = exp(sign*2*pi*i*f[j]*x[l]) built on the device from fp64 host coordinates."""
    f, fp = _darr(f)
    x, xp = _darr(x)
    E = torch.empty((len(f), len(x)), dtype=cdtype, device=dev)
    h, st = _ctx(E)
    h.check(lib.pb_mdft_basis(h.ptr, _CODE[cdtype], fp, len(f), xp, len(x), int(sign), _p(E), st))
    return E


def cgemm(A, B, opA=capi.OP_N, opB=capi.OP_N, alpha=1.0):
    m = A.shape[0] if opA in (capi.OP_N, capi.OP_C) else A.shape[1]
    k = A.shape[1] if opA in (capi.OP_N, capi.OP_C) else A.shape[0]
    n = B.shape[1] if opB in (capi.OP_N, capi.OP_C) else B.shape[0]
    kb = B.shape[0] if opB in (capi.OP_N, capi.OP_C) else B.shape[1]
    if k != kb:
        raise ValueError(f'matmul: inner dimensions {k} and {kb} differ')
    A, B = A.contiguous(), B.contiguous()
    out = torch.empty((m, n), dtype=A.dtype, device=A.device)
    h, st = _ctx(A)
    h.check(lib.pb_cgemm(h.ptr, _CODE[A.dtype], opA, opB, m, n, k, float(alpha), _p(A), A.shape[1], _p(B), B.shape[1],
                         _p(out), n, st))
    return out


def mdft_apply(Ey, Ex, a, norm, adjoint, left_first):
    my, ny = Ey.shape
    mx, nx = Ex.shape
    a = a.contiguous()
    want = (my, mx) if adjoint else (ny, nx)
    if tuple(a.shape) != want:
        raise ValueError(f'array of shape {tuple(a.shape)} does not match the executor ({want})')
    oshape = (ny, nx) if adjoint else (my, mx)
    out = torch.empty(oshape, dtype=a.dtype, device=a.device)
    nwork = lib.pb_mdft_work_elems(my, ny, mx, nx, int(adjoint), int(left_first))
    work = torch.empty(nwork, dtype=a.dtype, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_mdft_apply(h.ptr, _CODE[a.dtype], _p(Ey), _p(Ex), my, ny, mx, nx, _p(a), _p(out), float(norm),
                              int(adjoint), int(left_first), _p(work), st))
    return out


def mdft_tc_supported(my, ny, mx, nx):
    return bool(lib.pb_mdft_tc_supported(my, ny, mx, nx))


def mdft_tc_expand(E):
    """complex64 basis (m, n) -> TF32-split real expansions (hi, lo), each (2m, 2n) float32."""
    m, n = E.shape
    hi = torch.empty((2 * m, 2 * n), dtype=torch.float32, device=E.device)
    lo = torch.empty_like(hi)
    h, st = _ctx(E)
    h.check(lib.pb_mdft_tc_expand(h.ptr, _p(E), m, n, _p(hi), _p(lo), st))
    return hi, lo


def mdft_tc_apply(ex_hi, ex_lo, ey_hi, ey_lo, a, norm):
    my, ny = ey_hi.shape[0] // 2, ey_hi.shape[1] // 2
    mx, nx = ex_hi.shape[0] // 2, ex_hi.shape[1] // 2
    a = a.contiguous()
    if tuple(a.shape) != (ny, nx):
        raise ValueError(f'array of shape {tuple(a.shape)} does not match the executor ({(ny, nx)})')
    out = torch.empty((my, mx), dtype=torch.complex64, device=a.device)
    work = torch.empty(lib.pb_mdft_tc_work_bytes(my, ny, mx, nx), dtype=torch.uint8, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_mdft_tc_apply(h.ptr, _p(ex_hi), _p(ex_lo), _p(ey_hi), _p(ey_lo), my, ny, mx, nx, _p(a), _p(out),
                                 float(norm), _p(work), st))
    return out


def phase_screen(amp, opd, kscale):
    opd = opd.contiguous()
    if opd.dtype not in _CPLX_OF:
        opd = opd.to(torch.float32)
    if amp is None:
        amp_kind, amp_t = capi.AMP_NONE, None
    elif amp.dtype in (torch.bool, torch.uint8):
        amp_kind, amp_t = capi.AMP_U8, amp.contiguous()
    else:
        amp_kind, amp_t = capi.AMP_REAL, amp.to(opd.dtype).contiguous()
    out = torch.empty(opd.shape, dtype=_CPLX_OF[opd.dtype], device=opd.device)
    h, st = _ctx(opd)
    h.check(lib.pb_phase_screen(h.ptr, _CODE[opd.dtype], _p(amp_t), amp_kind, _p(opd), float(kscale), opd.numel(),
                                _p(out), st))
    return out


def intensity(field, weight=1.0, out=None):
    field = field.contiguous()
    acc = out is not None
    if out is None:
        out = torch.empty(field.shape, dtype=_REAL_OF[field.dtype], device=field.device)
    else:
        _check_out(out, field.shape, _REAL_OF[field.dtype], field.device)
        if not out.is_contiguous():
            raise ValueError('`out` must be contiguous')
    h, st = _ctx(field)
    h.check(lib.pb_intensity(h.ptr, _CODE[field.dtype], _p(field), field.numel(), float(weight), int(acc), _p(out), st))
    return out


_BINOPS = {'mul': 0, 'truediv': 1, 'add': 2, 'sub': 3}


def binary(op, a, b, reverse=False):
    """a (op) b for complex tensor a and tensor-or-scalar b (pb_binary)."""
    a = a.contiguous()
    if isinstance(b, torch.Tensor):
        if b.shape != a.shape:
            raise ValueError(f'shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}')
        b = b.to(a.dtype).contiguous()
        s = 0j
    else:
        s, b = complex(b), None
    out = torch.empty_like(a)
    h, st = _ctx(a)
    h.check(lib.pb_binary(h.ptr, _CODE[a.dtype], _BINOPS[op], _p(a), _p(b), s.real, s.imag, int(reverse), a.numel(),
                          _p(out), st))
    return out


def mul_outer(a, vy=None, vx=None, conj_y=False, conj_x=False, scale=1.0, out=None):
    """a[y, x] * vy[y] * vx[x] * scale; `a` may be a strided (sliced) view with unit column stride."""
    if a.stride(1) != 1:
        a = a.contiguous()
    ny, nx = a.shape
    if out is None:
        out = torch.empty((ny, nx), dtype=a.dtype, device=a.device)
    else:
        _check_out(out, (ny, nx), a.dtype, a.device)
    h, st = _ctx(a)
    h.check(lib.pb_mul_outer(h.ptr, _CODE[a.dtype], _p(a), a.stride(0), ny, nx, _p(vy), int(conj_y), _p(vx),
                             int(conj_x), float(scale), _p(out), out.stride(0), st))
    return out


def weighted_sum(modes, weights):
    modes = modes.contiguous()
    k = modes.shape[0]
    w, wp = _darr(weights)
    out = torch.empty(modes.shape[1:], dtype=modes.dtype, device=modes.device)
    h, st = _ctx(modes)
    h.check(lib.pb_weighted_sum(h.ptr, _CODE[modes.dtype], _p(modes), k, out.numel(), wp, _p(out), st))
    return out


def otf_normalize(D, which):
    D = D.contiguous()
    ny, nx = D.shape
    rd = _REAL_OF[D.dtype]
    mtf = torch.empty((ny, nx), dtype=rd, device=D.device) if which & 1 else None
    ptf = torch.empty((ny, nx), dtype=rd, device=D.device) if which & 2 else None
    otf = torch.empty((ny, nx), dtype=D.dtype, device=D.device) if which & 4 else None
    h, st = _ctx(D)
    h.check(lib.pb_otf_normalize(h.ptr, _CODE[D.dtype], _p(D), ny, nx, which, _p(mtf), _p(ptf), _p(otf), st))
    return mtf, ptf, otf


def encircled_energy(mtf, df, radii_mm):
    mtf = mtf.contiguous()
    ny, nx = mtf.shape
    r, rp = _darr(radii_mm)
    out = (C.c_double * len(r))()
    h, st = _ctx(mtf)
    h.check(lib.pb_encircled_energy(h.ptr, _CODE[mtf.dtype], _p(mtf), ny, nx, float(df), rp, len(r), out, st))
    return np.array(out[:])


def moments(data):
    data = data.contiguous()
    if data.dtype not in _CPLX_OF:
        data = data.to(torch.float32)
    ny, nx = data.shape
    sums = (C.c_double * 3)()
    h, st = _ctx(data)
    h.check(lib.pb_moments(h.ptr, _CODE[data.dtype], _p(data), ny, nx, sums, st))
    return sums[0], sums[1], sums[2]


def mask_multiply(a, m=None, *, b=None, w=None, conj=False, one_minus=False, real_out=False, scale=1.0, out=None):
    """out (+)= scale * (a - b) * f(m) * w in one pass (pb_mask_multiply).  `a`, `b` complex; `m` real or complex
    (f = conj(m) / 1 - m by flag); `w` real; real_out keeps the real part in a real tensor; a given `out` is
    accumulated into."""
    a = a.contiguous()
    rd = _REAL_OF[a.dtype]

    def same(t, name):
        if t is not None and tuple(t.shape) != tuple(a.shape):
            raise ValueError(f'shape mismatch: field {tuple(a.shape)} vs {name} {tuple(t.shape)}')
        return t
    kind = capi.MASK_REAL
    if m is not None:
        m = same(asdevice(m), 'mask')
        if m.is_complex():
            kind, m = capi.MASK_COMPLEX, m.to(a.dtype).contiguous()
        else:
            m = m.to(rd).contiguous()
    b = None if b is None else same(b, 'subtrahend').to(a.dtype).contiguous()
    w = None if w is None else same(asdevice(w), 'window').to(rd).contiguous()
    flags = (capi.MASK_CONJ if conj else 0) | (capi.MASK_ONE_MINUS if one_minus else 0) | \
            (capi.MASK_REAL_OUT if real_out else 0)
    if out is None:
        out = torch.empty(a.shape, dtype=rd if real_out else a.dtype, device=a.device)
    else:
        _check_out(out, a.shape, rd if real_out else a.dtype, a.device)
        if not out.is_contiguous():
            raise ValueError('`out` must be contiguous')
        flags |= capi.MASK_ACCUMULATE
    h, st = _ctx(a)
    h.check(lib.pb_mask_multiply(h.ptr, _CODE[a.dtype], _p(a), _p(b), _p(m), kind, flags, _p(w), float(scale),
                                 a.numel(), _p(out), st))
    return out


def field_adjoint(mode, field, bar, opd, kscale):
    """Adjoints of from_amp_and_phase (pb_field_adjoint): mode 0 phase (complex out), 1 / 2 amplitude (real out)."""
    bar = bar.contiguous()
    rd = _REAL_OF[bar.dtype]
    field = None if field is None else field.to(bar.dtype).contiguous()
    opd = None if opd is None else asdevice(opd).to(rd).contiguous()
    out = torch.empty(bar.shape, dtype=bar.dtype if mode == 0 else rd, device=bar.device)
    h, st = _ctx(bar)
    h.check(lib.pb_field_adjoint(h.ptr, _CODE[bar.dtype], int(mode), _p(field), _p(bar), _p(opd), float(kscale),
                                 bar.numel(), _p(out), st))
    return out


def component(which, field):
    """re / im / angle / abs of a complex tensor as a contiguous real tensor (pb_component)."""
    field = field.contiguous()
    out = torch.empty(field.shape, dtype=_REAL_OF[field.dtype], device=field.device)
    h, st = _ctx(field)
    h.check(lib.pb_component(h.ptr, _CODE[field.dtype], {'real': 0, 'imag': 1, 'angle': 2, 'abs': 3}[which], _p(field),
                             field.numel(), _p(out), st))
    return out


def dot(a, b, w=None):
    """sum w * a * conj(b) as a Python complex (pb_dot; fp64 accumulation)."""
    a = a.contiguous()
    b = b.to(a.dtype).contiguous()
    if a.shape != b.shape:
        raise ValueError(f'shape mismatch {tuple(a.shape)} vs {tuple(b.shape)}')
    w = None if w is None else asdevice(w).to(_REAL_OF[a.dtype]).contiguous()
    out = (C.c_double * 2)()
    h, st = _ctx(a)
    h.check(lib.pb_dot(h.ptr, _CODE[a.dtype], _p(a), _p(b), _p(w), a.numel(), out, st))
    return complex(out[0], out[1])


def mode_projection(modes, bar):
    """tensordot(modes(k, m, n), bar(m, n)) over the trailing axes (pb_mode_projection) -> host array (k,)."""
    modes = modes.contiguous()
    k = modes.shape[0]
    cplx = bar.is_complex()
    bar = bar.to(_CPLX_OF[modes.dtype] if cplx else modes.dtype).contiguous()
    if tuple(bar.shape) != tuple(modes.shape[1:]):
        raise ValueError(f'shape mismatch: modes {tuple(modes.shape)} vs gradient {tuple(bar.shape)}')
    out = (C.c_double * (2 * k))()
    h, st = _ctx(modes)
    h.check(lib.pb_mode_projection(h.ptr, _CODE[modes.dtype], _p(modes), k, bar.numel(), _p(bar), int(cplx), out, st))
    v = np.array(out[:]).reshape(k, 2)
    return v[:, 0] + 1j * v[:, 1] if cplx else v[:, 0].copy()


def otf_adjoint_seed(which, bar, D):
    """k-space gradient of mtf (1) / ptf (2) / otf (4) _from_psf_adjoint (pb_otf_adjoint_seed)."""
    D = D.contiguous()
    ny, nx = D.shape
    bar = asdevice(bar)
    bar = bar.to(D.dtype if which == 4 else _REAL_OF[D.dtype]).contiguous()
    if tuple(bar.shape) != (ny, nx):
        raise ValueError(f'shape mismatch: gradient {tuple(bar.shape)} vs transform {(ny, nx)}')
    out = torch.empty_like(D)
    h, st = _ctx(D)
    h.check(lib.pb_otf_adjoint_seed(h.ptr, _CODE[D.dtype], int(which), _p(bar), _p(D), ny, nx, _p(out), st))
    return out


def encircled_energy_adjoint_seed(shape, df, radii_mm, ee_bar, rdtype, dev):
    ny, nx = shape
    r, rp = _darr(radii_mm)
    b, bp = _darr(ee_bar)
    if len(r) != len(b):
        raise ValueError('one gradient value per radius is required')
    out = torch.empty((ny, nx), dtype=rdtype, device=dev)
    h, st = _ctx(out)
    h.check(lib.pb_encircled_energy_adjoint_seed(h.ptr, _CODE[rdtype], ny, nx, float(df), rp, bp, len(r), _p(out), st))
    return out


def vortex_phase(charge, xf, yf):
    xf = xf.contiguous()
    yf = yf.to(xf.dtype).contiguous()
    out = torch.empty(xf.shape, dtype=_CPLX_OF[xf.dtype], device=xf.device)
    h, st = _ctx(xf)
    h.check(lib.pb_vortex_phase(h.ptr, _CODE[xf.dtype], int(charge), _p(xf), _p(yf), xf.numel(), _p(out), st))
    return out


def resample_bilinear(cmap, xf, yf, cx, cy, dx, fill):
    """Bilinear, edge-replicating resampling of a complex map at focal coordinates; `fill` (tensor or scalar) outside."""
    xf = xf.contiguous()
    yf = yf.to(xf.dtype).contiguous()
    cd = _CPLX_OF[xf.dtype]
    cmap = cmap.to(cd).contiguous()
    ny, nx = cmap.shape
    fill_t, fs = None, 0j
    if isinstance(fill, torch.Tensor):
        fill_t = ascomplex(fill).to(cd).expand(xf.shape).contiguous()
    else:
        fs = complex(fill)
    out = torch.empty(xf.shape, dtype=cd, device=xf.device)
    h, st = _ctx(xf)
    h.check(lib.pb_resample_bilinear(h.ptr, _CODE[cd], _p(cmap), ny, nx, _p(xf), _p(yf), xf.numel(), float(cx), float(cy),
                                     float(dx), _p(fill_t), fs.real, fs.imag, _p(out), st))
    return out


def radial_window(shape, fdx, shift, here, nxt, rdtype, dev):
    """(window, xf, yf) of one multi-resolution level (pb_radial_window); here / nxt = (a, b) or None."""
    ny, nx = shape
    win, xf, yf = (torch.empty((ny, nx), dtype=rdtype, device=dev) for _ in range(3))
    a0, b0 = here if here is not None else (-1.0, 0.0)
    a1, b1 = nxt if nxt is not None else (-1.0, 0.0)
    h, st = _ctx(win)
    h.check(lib.pb_radial_window(h.ptr, _CODE[rdtype], ny, nx, float(fdx), float(shift), float(a0), float(b0), float(a1),
                                 float(b1), _p(win), _p(xf), _p(yf), st))
    return win, xf, yf


def _iarr(v):
    v = np.ascontiguousarray(v, dtype=np.int32)
    return v, v.ctypes.data_as(C.POINTER(C.c_int))


def xy_grid(shape, dx, rdtype, dev, want='xy'):
    """Any of x, y, r, t of make_xy_grid / cart_to_polar from one kernel (pb_xy_grid); returns a dict."""
    ny, nx = shape
    out = {k: torch.empty((ny, nx), dtype=rdtype, device=dev) for k in want}
    any_t = next(iter(out.values()))
    h, st = _ctx(any_t)
    h.check(lib.pb_xy_grid(h.ptr, _CODE[rdtype], ny, nx, float(dx), _p(out.get('x')), _p(out.get('y')), _p(out.get('r')),
                           _p(out.get('t')), st))
    return out


def cart_to_polar(x, y):
    x = x.contiguous()
    y = y.to(x.dtype).contiguous()
    r, t = torch.empty_like(x), torch.empty_like(x)
    h, st = _ctx(x)
    h.check(lib.pb_cart_to_polar(h.ptr, _CODE[x.dtype], _p(x), _p(y), x.numel(), _p(r), _p(t), st))
    return r, t


def circle(r, radius, aa_dx=0.0):
    """bool mask r - radius <= 0, or with aa_dx > 0 the one-sample grey-edge coverage (pb_circle)."""
    r = r.contiguous()
    out = torch.empty(r.shape, dtype=r.dtype if aa_dx > 0 else torch.bool, device=r.device)
    h, st = _ctx(r)
    h.check(lib.pb_circle(h.ptr, _CODE[r.dtype], _p(r), r.numel(), float(radius), float(aa_dx), _p(out), st))
    return out


def jacobi_seq(ns, alpha, beta, x):
    x = x.contiguous()
    ns = [int(n) for n in ns]
    nmax = max(ns) if ns else 0
    slots = np.full(nmax + 1, -1, dtype=np.int32)
    for i, n in enumerate(ns):
        slots[n] = i
    _, sp = _iarr(slots)
    out = torch.empty((len(ns),) + tuple(x.shape), dtype=x.dtype, device=x.device)
    if not ns:
        return out
    h, st = _ctx(x)
    h.check(lib.pb_jacobi_seq(h.ptr, _CODE[x.dtype], _p(x), x.numel(), nmax, float(alpha), float(beta), sp, _p(out), st))
    return out


def zernike_seq(nms, a, b, norm, polar):
    a = a.contiguous()
    b = b.to(a.dtype).contiguous()
    k = len(nms)
    out = torch.empty((k,) + tuple(a.shape), dtype=a.dtype, device=a.device)
    if k == 0:
        return out
    n, nptr = _iarr([nm[0] for nm in nms])
    m, mptr = _iarr([nm[1] for nm in nms])
    h, st = _ctx(a)
    h.check(lib.pb_zernike_seq(h.ptr, _CODE[a.dtype], int(polar), _p(a), _p(b), a.numel(), k, nptr, mptr, int(bool(norm)),
                               _p(out), st))
    return out


def zernike_sum(coefs, nms, a, b, norm, polar):
    a = a.contiguous()
    b = b.to(a.dtype).contiguous()
    k = len(nms)
    out = torch.empty(a.shape, dtype=a.dtype, device=a.device)
    n, nptr = _iarr([nm[0] for nm in nms])
    m, mptr = _iarr([nm[1] for nm in nms])
    c, cptr = _darr(np.asarray(coefs, dtype=np.float64)[:k])
    h, st = _ctx(a)
    h.check(lib.pb_zernike_sum(h.ptr, _CODE[a.dtype], int(polar), _p(a), _p(b), a.numel(), k, nptr, mptr, cptr,
                               int(bool(norm)), _p(out), st))
    return out


def balance_scale(a, b):
    """Device scalar s = sqrt(sum a^2 / sum b^2) (pb_balance_scale); stays on the device."""
    a = a.contiguous()
    b = b.to(a.dtype).contiguous()
    s = torch.empty(1, dtype=torch.float64, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_balance_scale(h.ptr, _CODE[a.dtype], _p(a), _p(b), a.numel(), _p(s), st))
    return s


def pack_complex(re, im=None, im_scale=None):
    re = re.contiguous()
    if re.dtype not in _CPLX_OF:
        re = re.to(torch.float32)
    im = None if im is None else im.to(re.dtype).contiguous()
    out = torch.empty(re.shape, dtype=_CPLX_OF[re.dtype], device=re.device)
    h, st = _ctx(re)
    h.check(lib.pb_pack_complex(h.ptr, _CODE[re.dtype], _p(re), _p(im), _p(im_scale), re.numel(), _p(out), st))
    return out


def packed_spectrum_product(Z, scale=1.0, im_scale=None):
    """O*H from Z = FFT2(o + i*s*h) of two real arrays (pb_packed_spectrum_product)."""
    Z = Z.contiguous()
    ny, nx = Z.shape
    out = torch.empty_like(Z)
    h, st = _ctx(Z)
    h.check(lib.pb_packed_spectrum_product(h.ptr, _CODE[Z.dtype], _p(Z), ny, nx, float(scale), _p(im_scale), _p(out), st))
    return out


def _real2d(a):
    a = asdevice(a)
    if a.dtype not in _CPLX_OF:
        a = a.to(torch.float32)
    return a.contiguous()


def bindown(a, fy, fx, mean):
    a = _real2d(a)
    ny, nx = a.shape
    out = torch.empty((ny // max(fy, 1), nx // max(fx, 1)), dtype=a.dtype, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_bindown(h.ptr, _CODE[a.dtype], _p(a), ny, nx, int(fy), int(fx), int(bool(mean)), _p(out), st))
    return out


def tile(a, fy, fx, scale):
    a = _real2d(a)
    ny, nx = a.shape
    out = torch.empty((ny * fy, nx * fx), dtype=a.dtype, device=a.device)
    h, st = _ctx(a)
    h.check(lib.pb_tile(h.ptr, _CODE[a.dtype], _p(a), ny, nx, int(fy), int(fx), float(scale), _p(out), st))
    return out


def separable_tf(kind, fx, fy, wx, wy):
    fx = _real2d(fx).reshape(-1)
    fy = _real2d(fy).to(fx.dtype).reshape(-1)
    out = torch.empty((fy.numel(), fx.numel()), dtype=fx.dtype, device=fx.device)
    h, st = _ctx(fx)
    h.check(lib.pb_separable_tf(h.ptr, _CODE[fx.dtype], int(kind), _p(fx), _p(fy), fy.numel(), fx.numel(), float(wx), float(wy),
                                _p(out), st))
    return out


def launch_count(dev=None):
    dev = device() if dev is None else torch.device(dev)
    return capi.launch_count(dev.index)


def ceil_half(d):
    return math.ceil(d / 2)
