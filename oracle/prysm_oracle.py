"""CPU oracle for the prysm propagation hot path -- TEST INFRASTRUCTURE ONLY.

This file is a numpy / scipy.fft restatement of the algorithms on the path named by
BASELINE.json (prysm.propagation + prysm.fttools + the psf/otf reductions, reference
v0.22 @ 5008f47).  It is the *checker* for the CUDA product in ``prysm_b200``: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu-baseline / reference arm
may import it.  Nothing in ``prysm_b200`` imports this module, and the product has no
CPU fallback.

Third-party arithmetic: like the reference, the FFT is ``scipy.fft`` (pocketfft) and
the dense contractions are numpy ``matmul`` (OpenBLAS); prysm pins no versions
(reference ``pyproject.toml:39-42``); this image has scipy 1.18.1 / numpy 2.3.5.

Parity pin: ``oracle/check_against_reference.py`` runs every function below against
the unmodified reference imported from /root/reference on seeded inputs (bit-for-bit or
<= 4 ulp fp64), and ``oracle/make_golden.py`` writes reference outputs to
``tests/golden/*.npz`` which ``tests/test_oracle_golden.py`` replays on any box.

Every function cites the reference lines it restates (paths relative to /root/reference).
Units follow the reference: wavelength um, pupil dx mm, focal dx um, OPD nm, efl / z mm.
Arrays are indexed [y, x].
"""
import math

import numpy as np
from scipy import fft as sfft

# --------------------------------------------------------------------------------------
# grids, pad, crop  (prysm/fttools.py:13-125)
# --------------------------------------------------------------------------------------


def fftrange(n, dtype=np.float64):
    """Integer grid whose zero sits at index n//2.  prysm/fttools.py:13-15."""
    lo = -(n // 2)
    return np.arange(lo, lo + n, dtype=dtype)


def next_fast_len(n):
    """5-smooth length used by CZT.  prysm/fttools.py:23-31 (-> scipy.fft.next_fast_len)."""
    return sfft.next_fast_len(n)


def fftfreq(n, d=1.0, dtype=np.float64):
    """Unshifted DFT sample frequencies cast to the working precision.  prysm/fttools.py:34-40."""
    return sfft.fftfreq(n, d).astype(dtype)


def padded_shape(shape, Q):
    """ceil(s*Q) per axis.  prysm/fttools.py:75."""
    return tuple(math.ceil(s * Q) for s in shape)


def pad_offsets(in_shape, out_shape):
    """Insertion offset ceil((out-in)/2) per axis.  prysm/fttools.py:87-88."""
    return tuple(math.ceil((o - i) / 2) for o, i in zip(out_shape, in_shape))


def pad2d(a, Q=2, value=0, out_shape=None):
    """Centred constant pad.  prysm/fttools.py:43-100 (mode='constant' branch)."""
    if Q == 1 and out_shape is None:
        return a
    if out_shape is None:
        out_shape = padded_shape(a.shape, Q)
    elif isinstance(out_shape, int):
        out_shape = (out_shape,) * a.ndim
    off = pad_offsets(a.shape, out_shape)
    out = np.full(out_shape, value, dtype=a.dtype) if value != 0 else np.zeros(out_shape, dtype=a.dtype)
    out[tuple(slice(o, o + s) for o, s in zip(off, a.shape))] = a
    return out


def crop_center(a, out_shape):
    """Inverse of pad2d: window starting at ceil((in-out)/2).  prysm/fttools.py:103-125."""
    if isinstance(out_shape, int):
        out_shape = (out_shape, out_shape)
    off = pad_offsets(out_shape, a.shape)
    return a[tuple(slice(o, o + s) for o, s in zip(off, out_shape))]


def shape_before_pad(shape, Q):
    """int(s // Q) per axis.  prysm/propagation/_kernels.py:14-18."""
    if Q == 1:
        return tuple(shape)
    return tuple(int(s // Q) for s in shape)


def adjoint_pad2d(a, Q):
    """prysm/propagation/_kernels.py:21-26."""
    tgt = shape_before_pad(a.shape, Q)
    return a if tgt == tuple(a.shape) else crop_center(a, tgt)


# --------------------------------------------------------------------------------------
# FFT focus / unfocus  (prysm/propagation/fft.py:7-85, 112-155)
# --------------------------------------------------------------------------------------


def _centered_ft2(a, inverse):
    f = sfft.ifft2 if inverse else sfft.fft2
    return sfft.fftshift(f(sfft.ifftshift(a), norm='ortho'))


def focus(w, Q):
    """pupil -> psf: fftshift(fft2(ifftshift(pad(w,Q)), ortho)).  prysm/propagation/fft.py:7-25."""
    return _centered_ft2(pad2d(w, Q) if Q != 1 else w, inverse=False)


def unfocus(w, Q):
    """psf -> pupil, same with ifft2.  prysm/propagation/fft.py:48-65."""
    return _centered_ft2(pad2d(w, Q) if Q != 1 else w, inverse=True)


def focus_adjoint(g, Q):
    """prysm/propagation/fft.py:28-45."""
    return adjoint_pad2d(_centered_ft2(g, inverse=True), Q)


def unfocus_adjoint(g, Q):
    """prysm/propagation/fft.py:68-85."""
    return adjoint_pad2d(_centered_ft2(g, inverse=False), Q)


def pupil_sample_to_psf_sample(pupil_dx, samples, wavelength, efl):
    """efl*wvl/(dx*K).  prysm/propagation/fft.py:112-132."""
    return (efl * wavelength) / (pupil_dx * samples)


def psf_sample_to_pupil_sample(psf_dx, samples, wavelength, efl):
    """prysm/propagation/fft.py:135-155."""
    return (efl * wavelength) / (psf_dx * samples)


# --------------------------------------------------------------------------------------
# wavefront synthesis and intensity  (prysm/propagation/wavefront.py:59-151, _kernels.py:40-43)
# --------------------------------------------------------------------------------------


def phase_prefix(wavelength):
    """i*2pi/wvl[um]/1e3: OPD[nm] -> radians.  prysm/propagation/_kernels.py:40-43."""
    return 1j * 2 * np.pi / wavelength / 1e3


def from_amp_and_phase(amp, opd_nm, wavelength):
    """A*exp(prefix*OPD).  prysm/propagation/wavefront.py:59-79."""
    if opd_nm is None:
        return amp
    return amp * np.exp(phase_prefix(wavelength) * opd_nm)


def phase_screen(opd_nm, wavelength):
    """prysm/propagation/wavefront.py:82-96."""
    return np.exp(phase_prefix(wavelength) * opd_nm)


def thin_lens(f, wavelength, x, y):
    """exp(-i*2pi/wvl_mm * r^2/(2f)).  prysm/propagation/wavefront.py:99-144."""
    w = wavelength / 1e3
    return np.exp((-1j * 2 * np.pi / w) * ((x * x + y * y) / (2 * f)))


def intensity(field):
    """re^2+im^2.  prysm/propagation/wavefront.py:147-151."""
    return field.real * field.real + field.imag * field.imag


# --------------------------------------------------------------------------------------
# angular spectrum  (prysm/propagation/angular_spectrum.py:9-114)
# --------------------------------------------------------------------------------------


def angular_spectrum_vectors(shape, wvl, dx, z, dtype=np.float64):
    """The two 1-D factors (tfy, tfx) of the separable Fresnel transfer function.

    prysm/propagation/angular_spectrum.py:102-113: wvl -> mm, k = fftfreq(s, dx).astype(P),
    exp(-i*pi*wvl*z*k^2) per axis.
    """
    if isinstance(shape, int):
        shape = (shape, shape)
    wmm = wvl / 1e3
    ky, kx = (sfft.fftfreq(s, dx).astype(dtype) for s in shape)
    pre = -1j * np.pi * wmm * z
    return np.exp(pre * (ky * ky)), np.exp(pre * (kx * kx))


def angular_spectrum_transfer_function(shape, wvl, dx, z, dtype=np.float64):
    """outer(tfy, tfx).  prysm/propagation/angular_spectrum.py:82-114."""
    ty, tx = angular_spectrum_vectors(shape, wvl, dx, z, dtype)
    return np.outer(ty, tx)


def angular_spectrum(field, wvl, dx, z, Q=2, tf=None, dtype=np.float64):
    """ifft2(fft2(pad(field)) * tf); result stays padded.  prysm/propagation/angular_spectrum.py:9-42."""
    if tf is not None:
        return sfft.ifft2(sfft.fft2(field) * tf)
    if Q != 1:
        field = pad2d(field, Q)
    tf = angular_spectrum_transfer_function(field.shape, wvl, dx, z, dtype)
    return sfft.ifft2(sfft.fft2(field) * tf)


def angular_spectrum_adjoint(g, wvl, dx, z, Q=2, tf=None, dtype=np.float64):
    """prysm/propagation/angular_spectrum.py:45-79."""
    if tf is None:
        tf = angular_spectrum_transfer_function(g.shape, wvl, dx, z, dtype)
        tgt = shape_before_pad(g.shape, Q)
    else:
        tgt = tuple(g.shape)
    out = sfft.ifft2(sfft.fft2(g) * np.conj(tf))
    return out if tgt == tuple(g.shape) else crop_center(out, tgt)


# --------------------------------------------------------------------------------------
# fixed-sampling executors  (prysm/fttools.py:155-535, prysm/propagation/dft.py:12-117)
# --------------------------------------------------------------------------------------


def coordinates_for_focus(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl,
                          focal_shift=(0, 0), dtype=np.float64):
    """x, y [mm] and fx, fy [1/mm].  prysm/propagation/dft.py:12-66."""
    if isinstance(pupil_samples, int):
        pupil_samples = (pupil_samples, pupil_samples)
    if isinstance(focal_samples, int):
        focal_samples = (focal_samples, focal_samples)
    pny, pnx = pupil_samples
    fny, fnx = focal_samples
    sx, sy = focal_shift
    x = fftrange(pnx, dtype) * pupil_dx
    y = fftrange(pny, dtype) * pupil_dx
    inv = 1.0 / (wavelength * efl)
    fx = (fftrange(fnx, dtype) * focal_dx + sx) * inv
    fy = (fftrange(fny, dtype) * focal_dx + sy) * inv
    return x, y, fx, fy


class MDFT:
    """Dense two-sided matrix DFT.  prysm/fttools.py:155-232."""

    def __init__(self, x, y, fx, fy, sign=-1, norm=1.0):
        c = sign * 2j * np.pi
        self.Ex = np.exp(c * np.outer(fx, x))
        self.Ey = np.exp(c * np.outer(fy, y))
        self.norm = norm
        Nx, Ny, Mx, My = len(x), len(y), len(fx), len(fy)
        # cheaper association first, prysm/fttools.py:198-199
        self._forward_left_first = My * Nx * (Ny + Mx) <= Ny * Mx * (Nx + My)
        self._adjoint_left_first = Ny * Mx * (My + Nx) <= My * Nx * (Mx + Ny)

    def __call__(self, a):
        if self._forward_left_first:
            return ((self.Ey @ a) @ self.Ex.T) * self.norm
        return (self.Ey @ (a @ self.Ex.T)) * self.norm

    def adjoint(self, g):
        EyH = self.Ey.conj().T
        ExC = self.Ex.conj()
        if self._adjoint_left_first:
            return ((EyH @ g) @ ExC) * self.norm
        return (EyH @ (g @ ExC)) * self.norm

    def nbytes(self):
        return self.Ex.nbytes + self.Ey.nbytes


def _czt_axis_basis(N, M, K, shift, alpha, rdtype, cdtype, sign):
    """Chirps a(m), b(n) and the kernel spectrum H.  prysm/fttools.py:372-389."""
    n = fftrange(N, rdtype)
    m = fftrange(M, rdtype)
    q = m + shift
    c = sign * 1j * np.pi * alpha
    a = np.exp(c * q * q)
    b = np.exp(c * n * n)
    d = np.arange(m[0] - n[-1], m[-1] - n[0] + 1, dtype=rdtype)
    h = np.zeros(K, dtype=cdtype)
    h[:len(d)] = np.exp(-c * (d + shift) * (d + shift))
    return sfft.fft(h), b, a


class CZT:
    """Bluestein chirp-z with the MDFT interface.  prysm/fttools.py:235-369."""

    def __init__(self, x, y, fx, fy, sign=-1, norm=1.0, rdtype=np.float64):
        if sign not in (-1, 1):
            raise ValueError(f'sign must be -1 or +1, got {sign}')
        cdtype = np.result_type(rdtype, 1j)
        self.sign, self.norm = sign, norm
        Nx, Mx, Ny, My = len(x), len(fx), len(y), len(fy)
        dx, dfx = float(x[1] - x[0]), float(fx[1] - fx[0])
        dy, dfy = float(y[1] - y[0]), float(fy[1] - fy[0])
        sh_x = float(fx[Mx // 2]) / dfx
        sh_y = float(fy[My // 2]) / dfy
        Kx, Ky = next_fast_len(Nx + Mx - 1), next_fast_len(Ny + My - 1)
        Hx, bx, ax = _czt_axis_basis(Nx, Mx, Kx, sh_x, dx * dfx, rdtype, cdtype, sign)
        Hy, by, ay = _czt_axis_basis(Ny, My, Ky, sh_y, dy * dfy, rdtype, cdtype, sign)
        c = sign * 2j * np.pi
        self.bx, self.Hx, self.ax = bx, Hx, ax
        self.by, self.Hy, self.ay = by[:, None], Hy[:, None], ay[:, None]
        self.px = np.exp(c * float(x[Nx // 2]) * fx)
        self.py = np.exp(c * float(y[Ny // 2]) * fy)[:, None]
        self.N, self.M, self.K = (Ny, Nx), (My, Mx), (Ky, Kx)
        cost_x_first = Ny * Kx * math.log2(Kx) + Mx * Ky * math.log2(Ky)
        cost_y_first = Nx * Ky * math.log2(Ky) + My * Kx * math.log2(Kx)
        self._x_first = cost_x_first <= cost_y_first

    def _along_x(self, o):
        (Ny, Nx), (My, Mx), (Ky, Kx) = self.N, self.M, self.K
        o = sfft.ifft(sfft.fft(o, Kx, axis=1) * self.Hx, axis=1)
        return o[:, Nx - 1:Nx - 1 + Mx] * self.ax * self.px

    def _along_y(self, o):
        (Ny, Nx), (My, Mx), (Ky, Kx) = self.N, self.M, self.K
        o = sfft.ifft(sfft.fft(o, Ky, axis=0) * self.Hy, axis=0)
        return o[Ny - 1:Ny - 1 + My] * self.ay * self.py

    def __call__(self, a):
        o = a * self.bx * self.by
        o = self._along_y(self._along_x(o)) if self._x_first else self._along_x(self._along_y(o))
        return o * self.norm

    def adjoint(self, g):
        (Ny, Nx), (My, Mx), (Ky, Kx) = self.N, self.M, self.K
        o = g * self.px.conj() * self.py.conj() * self.ax.conj() * self.ay.conj()

        def back_y(o):
            t = np.zeros((Ky, o.shape[1]), dtype=o.dtype)
            t[Ny - 1:Ny - 1 + My] = o
            return sfft.ifft(sfft.fft(t, axis=0) * self.Hy.conj(), axis=0)[:Ny]

        def back_x(o):
            t = np.zeros((o.shape[0], Kx), dtype=o.dtype)
            t[:, Nx - 1:Nx - 1 + Mx] = o
            return sfft.ifft(sfft.fft(t, axis=1) * self.Hx.conj(), axis=1)[:, :Nx]

        o = back_x(back_y(o)) if self._x_first else back_y(back_x(o))
        return o * self.bx.conj() * self.by.conj() * self.norm

    def nbytes(self):
        return sum(v.nbytes for v in (self.by, self.bx, self.Hy, self.Hx, self.ay, self.ax, self.px, self.py))


def _uniform_spacing(v, name, rdtype):
    """prysm/fttools.py:484-497."""
    if len(v) < 2:
        raise ValueError(f'{name} must contain at least two samples')
    sp = float(v[1] - v[0])
    if sp == 0:
        raise ValueError(f'{name} must have nonzero spacing')
    tol = 32 * np.finfo(rdtype).eps
    scale = max(1.0, abs(float(v[0])), abs(float(v[-1])), abs(sp))
    if not bool(np.allclose(np.diff(v), sp, rtol=tol, atol=tol * scale)):
        raise ValueError(f'{name} must be uniformly spaced')
    return sp


def _fft_compatible_length(alpha, N, M, name, rdtype):
    """prysm/fttools.py:500-514."""
    inv = 1 / abs(alpha)
    K = round(inv)
    tol = 32 * np.finfo(rdtype).eps
    if not math.isclose(inv, K, rel_tol=tol, abs_tol=tol):
        raise ValueError(f'{name} spacings are not FFT-compatible: '
                         'abs(input spacing * output spacing) must be 1/integer')
    if K < max(N, M):
        raise ValueError(f'{name} requires FFT length {K}, smaller than input/output length {max(N, M)}')
    return K


class FFTDFT:
    """One FFT per axis when dx*dfx = +-1/K.  prysm/fttools.py:392-481, 517-535."""

    def __init__(self, x, y, fx, fy, sign=-1, norm=1.0, rdtype=np.float64):
        if sign not in (-1, 1):
            raise ValueError(f'sign must be -1 or +1, got {sign}')
        Nx, Ny, Mx, My = len(x), len(y), len(fx), len(fy)
        dx, dy = _uniform_spacing(x, 'x', rdtype), _uniform_spacing(y, 'y', rdtype)
        dfx, dfy = _uniform_spacing(fx, 'fx', rdtype), _uniform_spacing(fy, 'fy', rdtype)
        Kx = _fft_compatible_length(dx * dfx, Nx, Mx, 'x/fx', rdtype)
        Ky = _fft_compatible_length(dy * dfy, Ny, My, 'y/fy', rdtype)
        c = sign * 2j * np.pi
        self.pre_x = np.exp(c * np.arange(Nx, dtype=rdtype) * dx * float(fx[0]))
        self.pre_y = np.exp(c * np.arange(Ny, dtype=rdtype) * dy * float(fy[0]))[:, None]
        self.post_x = np.exp(c * float(x[0]) * fx)
        self.post_y = np.exp(c * float(y[0]) * fy)[:, None]
        self.N, self.M, self.K = (Ny, Nx), (My, Mx), (Ky, Kx)
        self.dir_x = sign if dx * dfx > 0 else -sign
        self.dir_y = sign if dy * dfy > 0 else -sign
        self.norm = norm
        cost_x_first = Ny * Kx * math.log2(Kx) + Mx * Ky * math.log2(Ky)
        cost_y_first = Nx * Ky * math.log2(Ky) + My * Kx * math.log2(Kx)
        self._x_first = cost_x_first <= cost_y_first

    @staticmethod
    def _fwd(a, K, axis, d):
        return sfft.fft(a, K, axis=axis) if d == -1 else sfft.ifft(a, K, axis=axis) * K

    @staticmethod
    def _bwd(a, K, N, axis, d):
        shp = list(a.shape)
        shp[axis] = K
        t = np.zeros(shp, dtype=a.dtype)
        sl = [slice(None)] * a.ndim
        sl[axis] = slice(0, a.shape[axis])
        t[tuple(sl)] = a
        o = sfft.ifft(t, axis=axis) * K if d == -1 else sfft.fft(t, axis=axis)
        sl[axis] = slice(0, N)
        return o[tuple(sl)]

    def __call__(self, a):
        (Ny, Nx), (My, Mx), (Ky, Kx) = self.N, self.M, self.K
        o = a * self.pre_x * self.pre_y
        if self._x_first:
            o = self._fwd(o, Kx, 1, self.dir_x)[:, :Mx]
            o = self._fwd(o, Ky, 0, self.dir_y)[:My]
        else:
            o = self._fwd(o, Ky, 0, self.dir_y)[:My]
            o = self._fwd(o, Kx, 1, self.dir_x)[:, :Mx]
        return o * self.post_x * self.post_y * self.norm

    def adjoint(self, g):
        (Ny, Nx), (My, Mx), (Ky, Kx) = self.N, self.M, self.K
        o = g * self.post_x.conj() * self.post_y.conj()
        if self._x_first:
            o = self._bwd(o, Ky, Ny, 0, self.dir_y)
            o = self._bwd(o, Kx, Nx, 1, self.dir_x)
        else:
            o = self._bwd(o, Kx, Nx, 1, self.dir_x)
            o = self._bwd(o, Ky, Ny, 0, self.dir_y)
        return o * self.pre_x.conj() * self.pre_y.conj() * self.norm

    def nbytes(self):
        return sum(v.nbytes for v in (self.pre_x, self.pre_y, self.post_x, self.post_y))


def prepare_executor(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl,
                     focal_shift=(0, 0), kind='mdft', rdtype=np.float64):
    """prysm/propagation/dft.py:69-117: norm = pupil_dx*focal_dx/(wvl*efl) baked in."""
    x, y, fx, fy = coordinates_for_focus(pupil_dx, pupil_samples, focal_dx, focal_samples,
                                         wavelength, efl, focal_shift, rdtype)
    norm = (pupil_dx * focal_dx) / (wavelength * efl)
    if kind == 'mdft':
        op = MDFT(x, y, fx, fy, -1, norm)
    elif kind == 'czt':
        op = CZT(x, y, fx, fy, -1, norm, rdtype)
    elif kind == 'fftdft':
        op = FFTDFT(x, y, fx, fy, -1, norm, rdtype)
    else:
        raise ValueError(f"kind must be 'mdft', 'czt', or 'fftdft', got {kind!r}")
    op.pupil_dx, op.focal_dx = pupil_dx, focal_dx
    return op


def unit_cell_focal_grid(pupil_dx, pupil_diameter, wavelength, efl, Q=2):
    """prysm/propagation/dft.py:120-152."""
    m = math.ceil(Q * pupil_diameter / pupil_dx)
    return wavelength * efl / pupil_dx / m, m


def focus_dft(w, ex):
    """prysm/propagation/dft.py:297-313."""
    return ex(w)


def unfocus_dft(w, ex):
    """prysm/propagation/dft.py:335-351."""
    return ex.adjoint(w)


# --------------------------------------------------------------------------------------
# psf -> otf reductions  (prysm/otf.py:11-202), mode sum, centroid
# --------------------------------------------------------------------------------------


def transform_psf(psf, dx):
    """fftshift(fft2(ifftshift(psf))), df = 1000/(rows*dx).  prysm/otf.py:28-33."""
    data = sfft.fftshift(sfft.fft2(sfft.ifftshift(psf)))
    return data, 1000 / (data.shape[0] * dx)


def normalized_transform(psf, dx):
    """Divide by the sample at floor(s/2).  prysm/otf.py:11-13, 62-74."""
    data, df = transform_psf(psf, dx)
    cy, cx = (int(np.floor(s / 2)) for s in data.shape)
    return data / data[cy, cx], data, df


def mtf_from_psf(psf, dx):
    """prysm/otf.py:77-104."""
    n, _, df = normalized_transform(psf, dx)
    return abs(n), df


def ptf_from_psf(psf, dx):
    """prysm/otf.py:107-137."""
    n, _, df = normalized_transform(psf, dx)
    return np.angle(n), df


def otf_from_psf(psf, dx):
    """prysm/otf.py:140-167."""
    n, _, df = normalized_transform(psf, dx)
    return n, df


def encircled_energy(psf, dx, radius):
    """Baliga & Cohn: r * sum(MTF * J1(2 pi r nu)/nu) * dnu^2, nu on the fftrange*df grid with nu(0) -> 1e-16;
    radius in um -> mm.  prysm/otf.py:319-414."""
    from scipy.special import j1
    mtf, df = mtf_from_psf(psf, dx)
    fy = fftrange(mtf.shape[0]) * df
    fx = fftrange(mtf.shape[1]) * df
    nu = np.hypot(*np.meshgrid(fx, fy))
    nu[nu == 0] = 1e-16
    radii = np.atleast_1d(np.asarray(radius, dtype=np.float64)) / 1e3
    out = np.array([r * (mtf * j1(2 * np.pi * r * nu) / nu).sum() * df * df for r in radii])
    return float(out[0]) if np.ndim(radius) == 0 else out


def sum_of_2d_modes(modes, weights):
    """tensordot over the leading axis.  prysm/polynomials/fitting.py:7-37."""
    modes = np.asarray(modes)
    weights = np.asarray(weights).astype(modes.dtype)
    return np.tensordot(modes, weights, axes=(0, 0))


def centroid(data, dx=None, unit='spatial'):
    """Centre of mass; spatial = dx*(com - shape//2).  prysm/psf.py:174-203
    (scipy.ndimage.center_of_mass = sum(data*index)/sum(data) per axis)."""
    tot = data.sum()
    idx = np.indices(data.shape)
    com = tuple(float((data * g).sum() / tot) for g in idx)
    if unit != 'spatial':
        return com
    return tuple(dx * (c - s // 2) for c, s in zip(com, data.shape))


# --------------------------------------------------------------------------------------
# adjoint twins of the elementwise steps and of the otf reductions
# (prysm/propagation/_kernels.py:30-38, wavefront.py:172-298, otf.py:205-316, 417-471)
# --------------------------------------------------------------------------------------


def adjoint_multiply(grad, factor, real=False):
    """Adjoint w.r.t. x of y = x*factor: grad*conj(factor).  prysm/propagation/_kernels.py:30-38."""
    factor = np.asarray(factor)
    out = grad * (np.conj(factor) if np.iscomplexobj(factor) else factor)
    return np.real(out) if real else out


def intensity_adjoint(field, intensity_bar):
    """2 * Ibar * E.  prysm/propagation/wavefront.py:282-298."""
    return 2 * intensity_bar * field


def from_amp_and_phase_adjoint_phase(field, field_bar, wavelength):
    """prefix * imag(gbar * conj(g)) with the COMPLEX prefix i*2pi/(1e3*wvl) of the forward step
    (so the reference returns a purely imaginary array).  prysm/propagation/wavefront.py:172-188."""
    return phase_prefix(wavelength) * np.imag(field_bar * np.conj(field))


def from_amp_and_phase_adjoint_amp(field, field_bar, wavelength, phase=None):
    """real(gbar * conj(S)), S the unit phasor -- rebuilt from `phase`, else P/|P| (0 where P = 0).
    prysm/propagation/wavefront.py:190-225."""
    if phase is not None:
        S = np.exp(phase_prefix(wavelength) * phase)
        return np.real(field_bar * np.conj(S))
    mod = np.abs(field)
    ok = mod > 0
    g = np.real(field_bar * np.conj(field))
    return np.where(ok, g / np.where(ok, mod, 1), 0)


def thin_lens_adjoint(f, wavelength, x, y, lens_bar):
    """d/df of the thin-lens screen folded with its gradient: pi/(w f^2) * sum(r^2 imag(Lbar conj(L))).
    prysm/propagation/wavefront.py:244-280."""
    L = thin_lens(f, wavelength, x, y)
    w = wavelength / 1e3
    return np.pi / (w * f * f) * np.sum((x * x + y * y) * np.imag(lens_bar * np.conj(L)))


def transform_psf_adjoint(data_bar):
    """fftshift(ifft2(ifftshift(g), norm='forward')).  prysm/otf.py:36-59."""
    return sfft.fftshift(sfft.ifft2(sfft.ifftshift(data_bar), norm='forward'))


def _centre(shape):
    return tuple(int(np.floor(s / 2)) for s in shape)


def mtf_from_psf_adjoint(mtf_bar, data):
    """Through |.| and the division by the centre magnitude.  prysm/otf.py:205-242."""
    cy, cx = _centre(data.shape)
    mag = np.abs(data)
    a = mag[cy, cx]
    bar = mtf_bar * data / mag / a
    bar[cy, cx] -= np.sum(mtf_bar * mag) * data[cy, cx] / a ** 3
    return transform_psf_adjoint(bar).real


def ptf_from_psf_adjoint(ptf_bar, data):
    """Through angle(.) referenced to the centre phase.  prysm/otf.py:245-279."""
    cy, cx = _centre(data.shape)
    msq = data.real * data.real + data.imag * data.imag
    bar = ptf_bar * 1j * data / msq
    bar[cy, cx] -= np.sum(ptf_bar) * 1j * data[cy, cx] / msq[cy, cx]
    return transform_psf_adjoint(bar).real


def otf_from_psf_adjoint(otf_bar, data):
    """Through the division by the centre sample.  prysm/otf.py:282-316."""
    cy, cx = _centre(data.shape)
    cc = np.conj(data[cy, cx])
    bar = otf_bar / cc
    bar[cy, cx] -= np.sum(np.conj(data) * otf_bar) / cc ** 2
    return transform_psf_adjoint(bar).real


def encircled_energy_adjoint(ee_bar, data, dx, radius):
    """EE is linear in the MTF: fold the per-radius gradients into one MTF-plane gradient, then
    mtf_from_psf_adjoint.  prysm/otf.py:417-471 (geometry: otf.py:319-343)."""
    from scipy.special import j1
    df = 1000 / (data.shape[0] * dx)
    fy = fftrange(data.shape[0]) * df
    fx = fftrange(data.shape[1]) * df
    nu = np.hypot(*np.meshgrid(fx, fy))
    nu[nu == 0] = 1e-16
    radii = np.atleast_1d(np.asarray(radius, dtype=np.float64))
    bars = np.atleast_1d(np.asarray(ee_bar, dtype=np.float64))
    mtf_bar = 0.0
    for rb, r in zip(bars, radii):
        ri = r / 1e3
        mtf_bar = mtf_bar + rb * ri * (j1(2 * np.pi * ri * nu) / nu) * df * df
    return mtf_from_psf_adjoint(mtf_bar, data)


def sum_of_2d_modes_adjoint(modes, databar):
    """tensordot over the two trailing axes.  prysm/polynomials/fitting.py:40-57."""
    return np.tensordot(np.asarray(modes), databar)


# --------------------------------------------------------------------------------------
# Lyot-family coronagraph compositions  (prysm/propagation/coronagraph.py, dft.py:155-294)
# --------------------------------------------------------------------------------------


def focus_dft_adjoint(g, ex):
    """prysm/propagation/dft.py:316-332."""
    return ex.adjoint(g)


def unfocus_dft_adjoint(g, ex):
    """prysm/propagation/dft.py:354-370."""
    return ex(g)


def to_fpm_and_back(w, fpm, ex, return_more=False):
    """focus_dft -> * fpm -> unfocus_dft with one executor.  prysm/propagation/coronagraph.py:12-46."""
    at_fpm = focus_dft(w, ex)
    after_fpm = at_fpm * fpm
    nxt = unfocus_dft(after_fpm, ex)
    return (nxt, at_fpm, after_fpm) if return_more else nxt


def to_fpm_and_back_adjoint(g, fpm, ex, return_fpm_grad=False, field_at_fpm=None):
    """prysm/propagation/coronagraph.py:49-99.  Returns (Eabar, Ebbar, intermediate[, fpm_bar])."""
    if return_fpm_grad and field_at_fpm is None:
        raise ValueError('return_fpm_grad=True requires field_at_fpm from the forward propagation')
    Ebbar = unfocus_dft_adjoint(g, ex)
    inter = adjoint_multiply(Ebbar, fpm)
    Eabar = focus_dft_adjoint(inter, ex)
    if return_fpm_grad:
        return Eabar, Ebbar, inter, adjoint_multiply(Ebbar, field_at_fpm, real=not np.iscomplexobj(fpm))
    return Eabar, Ebbar, inter


def babinet(w, lyot, fpm, ex, return_more=False):
    """lyot * (w - to_fpm_and_back(w, 1 - fpm)).  prysm/propagation/coronagraph.py:301-353."""
    field, at_fpm, after_fpm = to_fpm_and_back(w, 1 - fpm, ex, return_more=True)
    at_lyot = w - field
    after_lyot = at_lyot if lyot is None else lyot * at_lyot
    return (after_lyot, at_fpm, after_fpm, at_lyot) if return_more else after_lyot


def babinet_adjoint(g, lyot, fpm, ex, field_at_fpm=None, field_at_lyot=None):
    """prysm/propagation/coronagraph.py:356-431.  Returns (abar, fpm_bar or None, lyot_bar or None); the
    gradients are produced when the matching forward fields are given."""
    lyot_complex = True if lyot is None else np.iscomplexobj(lyot)
    B = 1 - fpm
    cbar = g if lyot is None else adjoint_multiply(g, lyot)
    fpm_bar = None
    if field_at_fpm is not None:
        abar, _, _, fpm_bar = to_fpm_and_back_adjoint(cbar, B, ex, True, field_at_fpm)
    else:
        abar = to_fpm_and_back_adjoint(cbar, B, ex)[0]
    abar = cbar - abar
    lyot_bar = None if field_at_lyot is None else adjoint_multiply(g, field_at_lyot, real=not lyot_complex)
    return abar, fpm_bar, lyot_bar


def vortex_phase_mask(charge):
    """fpm(xf, yf) = exp(i*charge*atan2(yf, xf)).  prysm/propagation/coronagraph.py:102-132."""
    import numbers
    if not isinstance(charge, numbers.Integral):
        raise TypeError(f'charge must be an integer, got {charge!r}; non-integer charge has a branch cut at theta=pi')

    def fpm(xf, yf):
        return np.exp((1j * charge) * np.arctan2(yf, xf))
    return fpm


def prepare_measured_fpm(measurement, dx, center=(0, 0), charge=None, fill=None, order=1):
    """fpm(xf, yf): spline-interpolate a measured complex mask at col = (xf-cx)/dx + nx//2, row = (yf-cy)/dx + ny//2
    (scipy.ndimage.map_coordinates, mode='nearest'), `fill` (default: ideal vortex if charge is given, else 1) outside
    the measured extent.  prysm/propagation/coronagraph.py:135-209."""
    from scipy import ndimage
    meas = np.asarray(measurement)
    ny, nx = meas.shape
    cx, cy = center
    if fill is None:
        fill = vortex_phase_mask(charge) if charge is not None else 1.0

    def fpm(xf, yf):
        col = (xf - cx) / dx + nx // 2
        row = (yf - cy) / dx + ny // 2
        coords = np.stack([row.reshape(-1), col.reshape(-1)])
        ri = ndimage.map_coordinates(np.real(meas), coords, order=order, mode='nearest')
        ii = ndimage.map_coordinates(np.imag(meas), coords, order=order, mode='nearest')
        inside = (row >= 0) & (row <= ny - 1) & (col >= 0) & (col <= nx - 1)
        return np.where(inside, (ri + 1j * ii).reshape(xf.shape), fill(xf, yf) if callable(fill) else fill)
    return fpm


def smootherstep(t):
    """6t^5 - 15t^4 + 10t^3 on clip(t, 0, 1).  prysm/propagation/dft.py:155-158."""
    t = np.clip(t, 0, 1)
    return t * t * t * (t * (t * 6 - 15) + 10)


def cumulative_window(r, a, b):
    """1 inside a, 0 outside b, C2 in between.  prysm/propagation/dft.py:161-167."""
    return 1 - smootherstep((r - a) / (b - a))


class MultiResolutionExecutor:
    """prysm/propagation/dft.py:170-209."""

    def __init__(self, executors, windows, xf, yf):
        self.executors, self.windows, self.xf, self.yf = executors, windows, xf, yf

    def __len__(self):
        return len(self.executors)


def prepare_multiresolution(pupil_dx, pupil_samples, focal_dx, focal_samples, wavelength, efl, num_levels,
                            scaling=4.0, fine_samples=None, window=(0.2, 0.7), kind='mdft', rdtype=np.float64):
    """Level k: spacing focal_dx/scaling^k, half-sample focal shift, window = taper(own half-width) -
    taper(next level's half-width); the windows telescope to one.  prysm/propagation/dft.py:212-294."""
    if fine_samples is None:
        fine_samples = focal_samples
    inner, outer = window
    exs, xfs, yfs, radii, halves = [], [], [], [], []
    for k in range(num_levels):
        nf = focal_samples if k == 0 else fine_samples
        nfy, nfx = (nf, nf) if np.ndim(nf) == 0 else nf
        fdx = focal_dx / scaling ** k
        shift = fdx / 2.0
        exs.append(prepare_executor(pupil_dx, pupil_samples, fdx, (nfy, nfx), wavelength, efl,
                                    focal_shift=(shift, shift), kind=kind, rdtype=rdtype))
        xf, yf = np.meshgrid(fftrange(nfx, rdtype) * fdx + shift, fftrange(nfy, rdtype) * fdx + shift)
        xfs.append(xf)
        yfs.append(yf)
        radii.append(np.hypot(xf, yf))
        halves.append(min(nfy, nfx) / 2.0 * fdx)
    wins = []
    for k in range(num_levels):
        here = 1.0 if k == 0 else cumulative_window(radii[k], inner * halves[k], outer * halves[k])
        nxt = 0.0 if k == num_levels - 1 else cumulative_window(radii[k], inner * halves[k + 1], outer * halves[k + 1])
        wins.append(here - nxt)
    return MultiResolutionExecutor(exs, wins, xfs, yfs)


def to_fpm_and_back_multiresolution(w, fpm, mex, return_more=False):
    """Sum over levels of unfocus(focus(w) * fpm(xf, yf) * window).  prysm/propagation/coronagraph.py:212-251."""
    out, at, after = None, [], []
    for ex, win, xf, yf in zip(mex.executors, mex.windows, mex.xf, mex.yf):
        f = focus_dft(w, ex)
        g = f * fpm(xf, yf) * win
        c = unfocus_dft(g, ex)
        out = c if out is None else out + c
        at.append(f)
        after.append(g)
    return (out, at, after) if return_more else out


def to_fpm_and_back_multiresolution_adjoint(g, fpm, mex, field_at_fpm=None):
    """prysm/propagation/coronagraph.py:254-298.  Returns (Eabar, Ebbars, intermediates, fpm_bars or None)."""
    out, Ebbars, inters, bars = None, [], [], []
    for k, (ex, win, xf, yf) in enumerate(zip(mex.executors, mex.windows, mex.xf, mex.yf)):
        m = fpm(xf, yf)
        Eb = unfocus_dft_adjoint(g, ex)
        it = adjoint_multiply(Eb, m * win)
        c = focus_dft_adjoint(it, ex)
        out = c if out is None else out + c
        Ebbars.append(Eb)
        inters.append(it)
        if field_at_fpm is not None:
            bars.append(adjoint_multiply(Eb, field_at_fpm[k] * win, real=not np.iscomplexobj(m)))
    return out, Ebbars, inters, (bars if field_at_fpm is not None else None)


# --------------------------------------------------------------------------------------
# seeded synthetic pupil of SURVEY.md section 8(d)  (input generation; not on the hot path)
# --------------------------------------------------------------------------------------


def noll_to_nm(j):
    """Noll index -> (n, m), sign convention of prysm/polynomials/zernike.py:653-681
    (odd j -> negative m i.e. sine term)."""
    n = int(math.ceil((-1 + math.sqrt(1 + 8 * j)) / 2) - 1)
    if n == 0:
        return 0, 0
    k = j - (n * (n + 1)) // 2 - 1            # position within the order
    # order n holds |m| = n%2, then each following value twice, up to n
    seq = [n % 2] if n % 2 == 0 else [1, 1]
    while len(seq) < n + 1:
        seq += [seq[-1] + 2] * 2
    m_abs = seq[k]
    return n, (-m_abs if j % 2 else m_abs)


def zernike_nm(n, m, rho, theta):
    """Orthonormal (unit-RMS) Zernike: R_n^|m|(rho) * {cos, sin}(|m| theta) * norm,
    norm = sqrt(2(n+1)/(1+delta_m0)).  prysm/polynomials/zernike.py:25-27, 35-71."""
    am = abs(m)
    R = np.zeros_like(rho)
    for s in range((n - am) // 2 + 1):
        c = ((-1) ** s * math.factorial(n - s)
             / (math.factorial(s) * math.factorial((n + am) // 2 - s) * math.factorial((n - am) // 2 - s)))
        R = R + c * rho ** (n - 2 * s)
    if m < 0:
        R = R * np.sin(am * theta)
    elif m > 0:
        R = R * np.cos(am * theta)
    return R * math.sqrt(2 * (n + 1) / (2 if m == 0 else 1))


# --------------------------------------------------------------------------------------
# image-chain consumers of the same FFTs  (prysm/convolution.py:9-114, fttools.py:538-593)
# --------------------------------------------------------------------------------------


def conv(obj, psf):
    """fftshift(ifft2(fft2(ifftshift(o)) * fft2(ifftshift(h)))), real part for a real object.  prysm/convolution.py:9-32."""
    O = sfft.fft2(sfft.ifftshift(obj))
    H = sfft.fft2(sfft.ifftshift(psf))
    i = sfft.fftshift(sfft.ifft2(O * H))
    return i.real if not np.iscomplexobj(obj) else i


def forward_ft_unit(dx, samples, shift=True, dtype=np.float64):
    """fftfreq(samples, dx), fftshifted on request.  prysm/fttools.py:128-152."""
    u = fftfreq(samples, dx, dtype)
    return sfft.fftshift(u) if shift else u


def transfer_function_grids(shape, dx, shift=False, dtype=np.float64):
    """fx (1, N), fy (M, 1), fr, ft handed to callable transfer functions.  prysm/convolution.py:73-85."""
    uy, ux = (forward_ft_unit(dx, n, shift, dtype) for n in shape)
    fx, fy = ux.reshape(1, -1), uy.reshape(-1, 1)
    return fx, fy, np.hypot(fx, fy), np.arctan2(fy, fx)


def apply_transfer_functions(obj, tfs, shift=False):
    """Object spectrum times each transfer-function ARRAY, back to the image.  prysm/convolution.py:35-114 (callables are
    evaluated by the caller on transfer_function_grids)."""
    O = sfft.fft2(sfft.ifftshift(obj))
    if shift:
        O = sfft.fftshift(O)
    for tf in tfs:
        O = O * tf
    i = sfft.fftshift(sfft.ifft2(sfft.ifftshift(O) if shift else O))
    return i.real if not np.iscomplexobj(obj) else i


def fourier_resample(f, zoom, rdtype=np.float64):
    """Centered FFT, then a sign=+1 matrix DFT onto M = int(m*zoom) samples spaced 1/zoom, / (m n).
    prysm/fttools.py:538-593."""
    if zoom == 1:
        return f
    zoom = (zoom, zoom) if isinstance(zoom, (float, int)) else tuple(float(z) for z in zoom)
    if len(zoom) != 2 or any(z <= 0 for z in zoom):
        raise ValueError('zoom must contain two positive values')
    m, n = f.shape
    M, N = int(m * zoom[0]), int(n * zoom[1])
    if M < 1 or N < 1:
        raise ValueError('zoom produces an empty output')
    F = sfft.fftshift(sfft.fft2(sfft.ifftshift(f)))
    x, y = fftrange(n, rdtype), fftrange(m, rdtype)
    fx = fftrange(N, rdtype) * (1.0 / zoom[1] / n)
    fy = fftrange(M, rdtype) * (1.0 / zoom[0] / m)
    out = MDFT(x, y, fx, fy, +1)(F) * (1 / (m * n))
    return out if np.iscomplexobj(f) else out.real


def bindown(array, factor, mode='avg'):
    """Block mean / sum over factor x factor samples (2-D).  prysm/detector.py:222-274."""
    fy, fx = (factor, factor) if np.ndim(factor) == 0 else factor
    m, n = array.shape
    view = array.reshape(m // fy, fy, n // fx, fx)
    if mode.lower() in ('avg', 'average', 'mean'):
        return view.mean(axis=(1, 3))
    if mode.lower() == 'sum':
        return view.sum(axis=(1, 3))
    raise ValueError('mode must be average or sum.')


def tile(array, factor, scaling='sum'):
    """Repeat every sample factor x factor times (the adjoint of bindown), scaled by 1/prod(factor) for 'sum'.
    prysm/detector.py:277-338."""
    fy, fx = (factor, factor) if np.ndim(factor) == 0 else factor
    if scaling == 'sum':
        sf = 1 / (fy * fx)
    elif scaling in ('avg', 'average', 'mean'):
        sf = 1
    else:
        raise ValueError('scaling must be average or sum')
    return np.repeat(np.repeat(array, fy, axis=0), fx, axis=1) * sf


def pixel_ft(fx, fy, width_x, width_y):
    """sinc(fx wx) sinc(fy wy).  prysm/detector.py:174-194."""
    return np.sinc(fx * width_x) * np.sinc(fy * width_y)


def olpf_ft(fx, fy, width_x, width_y):
    """cos(2 wx fx) cos(2 wy fy).  prysm/detector.py:151-171."""
    return np.cos(2 * width_x * fx) * np.cos(2 * width_y * fy)


# --------------------------------------------------------------------------------------
# pupil synthesis by recurrence -- the step before the path
# (prysm/coordinates.py:73-102, 344-378; geometry.py:11-34, 337-372; polynomials/jacobi.py:13-175;
#  polynomials/_recurrence.py:9-57; polynomials/zernike.py:25-181, 633-690)
# --------------------------------------------------------------------------------------


def make_xy_grid(shape, dx=0, diameter=0, grid=True, dtype=np.float64):
    """x, y = fftrange(s)*dx per axis, meshgridded.  prysm/coordinates.py:344-378."""
    if not isinstance(shape, tuple):
        shape = (shape, shape)
    if diameter != 0:
        dx = diameter / max(shape)
    y, x = (fftrange(s, dtype) * dx for s in shape)
    if grid:
        x, y = np.meshgrid(x, y)
    return x, y


def cart_to_polar(x, y):
    """rho = hypot(x, y), phi = arctan2(y, x).  prysm/coordinates.py:73-102."""
    return np.hypot(x, y), np.arctan2(y, x)


def circle(radius, r):
    """r - radius <= 0.  prysm/geometry.py:337-372."""
    return (r - radius) <= 0


def antialias(d, dx):
    """Signed distance -> coverage with a one-sample ramp: clip(0.5 - d/dx, 0, 1).  prysm/geometry.py:11-34."""
    return np.minimum(np.maximum(0.5 - d / dx, 0), 1)


def recurrence_abc(n, alpha, beta):
    """DLMF 18.9 three-term coefficients, written as the reference writes them.  prysm/polynomials/jacobi.py:13-39."""
    aplusb = alpha + beta
    if n == 0 and (aplusb == 0 or aplusb == -1):
        return 1 / 2 * (alpha + beta) + 1, 1 / 2 * (alpha - beta), 1
    A = ((2 * n + alpha + beta + 1) * (2 * n + alpha + beta + 2)) / (2 * (n + 1) * (n + alpha + beta + 1))
    B = ((alpha ** 2 - beta ** 2) * (2 * n + alpha + beta + 1)) / (2 * (n + 1) * (n + alpha + beta + 1) * (2 * n + alpha + beta))
    C = ((n + alpha) * (n + beta) * (2 * n + alpha + beta + 2)) / ((n + 1) * (n + alpha + beta + 1) * (2 * n + alpha + beta))
    return A, B, C


def jacobi_seq(ns, alpha, beta, x):
    """P_n^(alpha,beta)(x) for ascending orders ns by P_k = (A x + B) P_{k-1} - C P_{k-2}, coefficients of order
    k-1; P_0 = 1, P_1 = alpha + 1 + (alpha + beta + 2)(x - 1)/2.  prysm/polynomials/jacobi.py:147-175,
    _recurrence.py:9-57."""
    ns = list(ns)
    out = np.empty((len(ns),) + x.shape, dtype=x.dtype)
    Pm2 = np.ones_like(x)
    Pm1 = alpha + 1 + (alpha + beta + 2) * ((x - 1) / 2)
    want = {n: i for i, n in enumerate(ns)}
    if 0 in want:
        out[want[0]] = Pm2
    if 1 in want:
        out[want[1]] = Pm1
    for k in range(2, (ns[-1] if ns else 0) + 1):
        A, B, C = recurrence_abc(k - 1, alpha, beta)
        Pn = (A * x + B) * Pm1 - C * Pm2
        Pm2, Pm1 = Pm1, Pn
        if k in want:
            out[want[k]] = Pn
    return out


def jacobi(n, alpha, beta, x):
    """prysm/polynomials/jacobi.py:42-79."""
    return jacobi_seq([n], alpha, beta, x)[0]


def zernike_norm(n, m):
    """sqrt(2(n+1)/(1+delta_m0)).  prysm/polynomials/zernike.py:25-27."""
    return math.sqrt((2 * (n + 1)) / (1 + (1 if m == 0 else 0)))


def zernike_nm_seq(nms, r, t, norm=True):
    """Z_n^m = norm * P_{(n-|m|)/2}^{(0,|m|)}(2r^2 - 1) * r^|m| * {cos(m t) | sin(|m| t)} with the Jacobi sequences
    shared per |m|.  prysm/polynomials/zernike.py:74-166."""
    x = 2 * (r * r) - 1
    nj_max = {}
    for n, m in nms:
        am = abs(m)
        nj_max[am] = max(nj_max.get(am, 0), (n - am) // 2)
    seqs = {am: jacobi_seq(range(nj + 1), 0, am, x) for am, nj in nj_max.items()}
    out = np.empty((len(nms),) + r.shape, dtype=r.dtype)
    for k, (n, m) in enumerate(nms):
        am = abs(m)
        jac = seqs[am][(n - am) // 2]
        if norm:
            jac = jac * zernike_norm(n, m)
        if m == 0:
            out[k] = jac
        else:
            out[k] = jac * (np.sin(am * t) if m < 0 else np.cos(am * t)) * r ** am
    return out


def zernike_sum(coefs, nms, x, y, norm=True):
    """sum_k c_k Z_k on Cartesian coordinates (zero coefficients skipped).  prysm/polynomials/zernike.py:169-181."""
    nms = tuple(nms)
    if not nms:
        return np.zeros_like(x)
    r, t = cart_to_polar(x, y)
    Z = zernike_nm_seq(nms, r, t, norm)
    z = np.zeros_like(x)
    for c, Zi in zip(coefs, Z):
        if c != 0.0:
            z = z + c * Zi
    return z


def synthetic_pupil(N, rdtype=np.float32, seed=20260923, sigma_nm=30.0, diameter=10.0, nmodes=36):
    """SURVEY.md section 8(d) builder: circular aperture of `diameter` mm on an N x N grid
    (dx = diameter/N, prysm/coordinates.py:344-378), OPD = sum of Noll 2..nmodes+1 orthonormal
    Zernikes with seeded N(0, sigma) nm coefficients.  Returns amp(bool), opd(rdtype, nm), dx."""
    dx = diameter / N
    g = fftrange(N, np.float64) * dx
    x, y = np.meshgrid(g, g)
    r = np.sqrt(x * x + y * y)                # prysm/coordinates.py cart_to_polar
    t = np.arctan2(y, x)
    amp = (r - diameter / 2) <= 0             # prysm/geometry.py:356-372
    rho = r / (diameter / 2)
    coefs = np.random.default_rng(seed).normal(0, sigma_nm, nmodes)
    opd = np.zeros((N, N))
    for j, c in zip(range(2, 2 + nmodes), coefs):
        n, m = noll_to_nm(j)
        opd += c * zernike_nm(n, m, rho, t)
    return amp, opd.astype(rdtype), dx
