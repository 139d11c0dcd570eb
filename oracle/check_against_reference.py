"""Pin the oracle against the UNMODIFIED reference (runs only where /root/reference exists).

    PYTHONDONTWRITEBYTECODE=1 python oracle/check_against_reference.py

Each oracle function is run beside the reference's own function on identical seeded inputs
(fp64 and fp32 config) and must agree to a few ulp (identical library calls -> usually
bit-identical).  Exit status 0 = oracle pinned.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
REF = os.environ.get('PRYSM_REFERENCE', '/root/reference')
sys.path.insert(0, REF)
sys.dont_write_bytecode = True

import prysm_oracle as O  # noqa: E402
from prysm.conf import config  # noqa: E402
from prysm import fttools, propagation, otf, psf  # noqa: E402
from prysm.propagation import Wavefront  # noqa: E402
from prysm.coordinates import make_xy_grid, cart_to_polar  # noqa: E402
from prysm.geometry import circle  # noqa: E402
from prysm.polynomials import zernike_nm_seq, noll_to_nm, sum_of_2d_modes  # noqa: E402

fails = []


def check(name, a, b, tol=0.0):
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype == bool:
        a, b = a.astype(np.int8), np.asarray(b).astype(np.int8)
    if a.shape != b.shape:
        fails.append(f'{name}: shape {a.shape} vs {b.shape}')
        print('FAIL', fails[-1])
        return
    den = max(float(np.abs(b).max()), 1e-300)
    err = float(np.abs(a - b).max()) / den
    ok = err <= tol
    print(('ok  ' if ok else 'FAIL'), f'{name:48s} rel_linf={err:.2e} (tol {tol:.0e})')
    if not ok:
        fails.append(name)


def crand(rng, shape, dt):
    return (rng.standard_normal(shape) + 1j * rng.standard_normal(shape)).astype(dt)


for prec, rdt, cdt, eps in ((64, np.float64, np.complex128, 1e-13), (32, np.float32, np.complex64, 2e-6)):
    config.precision = prec
    rng = np.random.default_rng(7)
    tag = f'[p{prec}] '
    # pad / crop incl. odd sizes and fractional Q
    for shp in ((8, 8), (9, 12), (7, 9), (12, 5)):
        a = crand(rng, shp, cdt)
        for Q in (1, 1.5, 2, 3):
            check(tag + f'pad2d {shp} Q={Q}', O.pad2d(a, Q), fttools.pad2d(a, Q))
            check(tag + f'focus {shp} Q={Q}', O.focus(a, Q), propagation.focus(a, Q))
            check(tag + f'unfocus {shp} Q={Q}', O.unfocus(a, Q), propagation.unfocus(a, Q))
            b = crand(rng, O.padded_shape(shp, Q), cdt)
            check(tag + f'focus_adjoint {shp} Q={Q}', O.focus_adjoint(b, Q), propagation.focus_adjoint(b, Q))
            check(tag + f'unfocus_adjoint {shp} Q={Q}', O.unfocus_adjoint(b, Q), propagation.unfocus_adjoint(b, Q))
        check(tag + f'crop_center {shp}', O.crop_center(O.pad2d(a, 2), shp), fttools.crop_center(fttools.pad2d(a, 2), shp))
    # wavefront synthesis
    amp = rng.random((16, 20)) > 0.3
    opd = (rng.standard_normal((16, 20)) * 100).astype(rdt)
    wf = Wavefront.from_amp_and_phase(amp, opd, 0.6328, 0.1)
    check(tag + 'from_amp_and_phase', O.from_amp_and_phase(amp, opd, 0.6328), wf.data)
    check(tag + 'phase_screen', O.phase_screen(opd, 0.55), Wavefront.phase_screen(opd, 0.55, 0.1).data)
    check(tag + 'intensity', O.intensity(wf.data), wf.intensity.data)
    x, y = make_xy_grid(16, diameter=10.0)
    check(tag + 'thin_lens', O.thin_lens(250.0, 0.6328, x, y), Wavefront.thin_lens(250.0, 0.6328, x, y).data)
    # angular spectrum
    f = crand(rng, (12, 16), cdt)
    for Q in (1, 2):
        check(tag + f'angular_spectrum Q={Q}', O.angular_spectrum(f, 0.6328, 0.05, 12.5, Q, dtype=rdt),
              propagation.angular_spectrum(f, 0.6328, 0.05, 12.5, Q))
    tf = propagation.angular_spectrum_transfer_function((12, 16), 0.6328, 0.05, 12.5)
    check(tag + 'as transfer function', O.angular_spectrum_transfer_function((12, 16), 0.6328, 0.05, 12.5, rdt), tf)
    check(tag + 'angular_spectrum tf', O.angular_spectrum(f, 0, 0, 0, tf=tf), propagation.angular_spectrum(f, 0, 0, 0, tf=tf))
    g = crand(rng, (24, 32), cdt)
    check(tag + 'angular_spectrum_adjoint', O.angular_spectrum_adjoint(g, 0.6328, 0.05, 12.5, 2, dtype=rdt),
          propagation.angular_spectrum_adjoint(g, 0.6328, 0.05, 12.5, 2))
    # executors
    for (pn, fn, fdx, shift) in (((16, 16), (8, 8), 2.0, (0, 0)), ((9, 12), (8, 11), 1.7, (3.0, -2.0)), ((32, 24), (40, 12), 0.9, (0.5, 0.25))):
        a = crand(rng, pn, cdt)
        for kind in ('mdft', 'czt'):
            r = propagation.prepare_executor(0.1, pn, fdx, fn, 0.6328, 100.0, shift, kind)
            o = O.prepare_executor(0.1, pn, fdx, fn, 0.6328, 100.0, shift, kind, rdt)
            check(tag + f'{kind} fwd {pn}->{fn}', o(a), r(a), eps)
            gg = crand(rng, fn, cdt)
            check(tag + f'{kind} adj {pn}->{fn}', o.adjoint(gg), r.adjoint(gg), eps)
            assert o.nbytes() == r.nbytes(), (kind, o.nbytes(), r.nbytes())
    if prec == 64:  # fftdft-compatible grid (float64 spacing test)
        K = 32
        pdx = 0.1
        fdx = 0.6328 * 100.0 / (pdx * K)
        for pn, fn in (((16, 16), (32, 32)), ((20, 16), (12, 32))):
            a = crand(rng, pn, cdt)
            r = propagation.prepare_executor(pdx, pn, fdx, fn, 0.6328, 100.0, (0, 0), 'fftdft')
            o = O.prepare_executor(pdx, pn, fdx, fn, 0.6328, 100.0, (0, 0), 'fftdft', rdt)
            check(tag + f'fftdft fwd {pn}->{fn}', o(a), r(a), eps)
            gg = crand(rng, fn, cdt)
            check(tag + f'fftdft adj {pn}->{fn}', o.adjoint(gg), r.adjoint(gg), eps)
    # otf
    p = rng.random((16, 16)).astype(rdt)
    check(tag + 'transform_psf', O.transform_psf(p, 1.5)[0], otf.transform_psf(p, 1.5)[0])
    check(tag + 'mtf', O.mtf_from_psf(p, 1.5)[0], otf.mtf_from_psf(p, 1.5).data)
    check(tag + 'ptf', O.ptf_from_psf(p, 1.5)[0], otf.ptf_from_psf(p, 1.5).data)
    check(tag + 'otf', O.otf_from_psf(p, 1.5)[0], otf.otf_from_psf(p, 1.5).data)
    assert O.mtf_from_psf(p, 1.5)[1] == otf.mtf_from_psf(p, 1.5).dx
    check(tag + 'encircled_energy', O.encircled_energy(p, 1.5, [2.0, 7.5]), otf.encircled_energy(p, 1.5, [2.0, 7.5]), 1e-6 if prec == 32 else 1e-13)
    # reductions
    modes = rng.random((5, 8, 9)).astype(rdt)
    wts = rng.random(5).astype(rdt)
    check(tag + 'sum_of_2d_modes', O.sum_of_2d_modes(modes, wts), sum_of_2d_modes(modes, wts))
    check(tag + 'centroid spatial', O.centroid(p.astype(np.float64), 1.5), psf.centroid(p.astype(np.float64), 1.5), 1e-12)
    check(tag + 'centroid pixels', O.centroid(p.astype(np.float64), unit='pixels'), psf.centroid(p.astype(np.float64), unit='pixels'), 1e-12)

# pupil builder (fp64 maths, cast at the end)
for j in range(1, 60):
    assert O.noll_to_nm(j) == noll_to_nm(j), j
config.precision = 64
N = 128
x, y = make_xy_grid(N, diameter=10.0)
r, t = cart_to_polar(x, y)
amp_ref = circle(5.0, r)
nms = [noll_to_nm(j) for j in range(2, 38)]
coefs = np.random.default_rng(20260923).normal(0, 30.0, 36)
opd_ref = sum_of_2d_modes(zernike_nm_seq(nms, r / 5.0, t), coefs)
amp, opd, dx = O.synthetic_pupil(N, np.float64)
check('synthetic_pupil amp', amp, amp_ref)
check('synthetic_pupil opd (inside aperture)', opd * amp, opd_ref * amp_ref, 1e-10)
assert dx == 10.0 / N

print()
if fails:
    print(f'{len(fails)} FAILURES:', fails)
    sys.exit(1)
print('oracle pinned against the reference: all checks passed')
