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

    # ---- adjoint twins of the elementwise steps and the otf reductions
    from prysm.polynomials import sum_of_2d_modes_adjoint
    a = rng.random((9, 12)).astype(rdt)
    ph = (rng.standard_normal((9, 12)) * 40).astype(rdt)
    a[2, 3] = 0
    wf = Wavefront.from_amp_and_phase(a, ph, 0.55, 0.1)
    bar = Wavefront(crand(rng, (9, 12), cdt), 0.55, 0.1)
    ibar = rng.random((9, 12)).astype(rdt)
    check(tag + 'intensity_adjoint', O.intensity_adjoint(wf.data, ibar), wf.intensity_adjoint(ibar).data, eps)
    check(tag + 'from_amp_and_phase_adjoint_phase', O.from_amp_and_phase_adjoint_phase(wf.data, bar.data, 0.55),
          wf.from_amp_and_phase_adjoint_phase(bar), eps)
    check(tag + 'from_amp_and_phase_adjoint_amp', O.from_amp_and_phase_adjoint_amp(wf.data, bar.data, 0.55),
          wf.from_amp_and_phase_adjoint_amp(bar), eps)
    check(tag + 'from_amp_and_phase_adjoint_amp(phase)', O.from_amp_and_phase_adjoint_amp(wf.data, bar.data, 0.55, ph),
          wf.from_amp_and_phase_adjoint_amp(bar, phase=ph), eps)
    xg, yg = make_xy_grid((9, 12), dx=0.1)
    xg, yg = xg.astype(rdt), yg.astype(rdt)
    check(tag + 'thin_lens_adjoint', O.thin_lens_adjoint(250.0, 0.55, xg, yg, bar.data),
          Wavefront.thin_lens_adjoint(250.0, 0.55, xg, yg, bar), 1e-5 if prec == 32 else 1e-12)
    pp = rng.random((10, 13)).astype(rdt)
    D = otf.transform_psf(pp, 1.5)[0]
    rb = rng.standard_normal((10, 13)).astype(rdt)
    cb = crand(rng, (10, 13), cdt)
    check(tag + 'mtf_from_psf_adjoint', O.mtf_from_psf_adjoint(rb, D), otf.mtf_from_psf_adjoint(rb, data=D), eps)
    check(tag + 'ptf_from_psf_adjoint', O.ptf_from_psf_adjoint(rb, D), otf.ptf_from_psf_adjoint(rb, data=D), eps)
    check(tag + 'otf_from_psf_adjoint', O.otf_from_psf_adjoint(cb, D), otf.otf_from_psf_adjoint(cb, data=D), eps)
    check(tag + 'encircled_energy_adjoint', O.encircled_energy_adjoint([0.3, -1.2], D, 1.5, [2.0, 7.5]),
          otf.encircled_energy_adjoint([0.3, -1.2], dx=1.5, radius=[2.0, 7.5], data=D), 1e-5 if prec == 32 else 1e-12)
    check(tag + 'sum_of_2d_modes_adjoint', O.sum_of_2d_modes_adjoint(modes, pp[:8, :9]), sum_of_2d_modes_adjoint(modes, pp[:8, :9]), eps)
    # ---- coronagraph compositions (MDFT / CZT executors)
    w = crand(rng, (14, 12), cdt)
    g = crand(rng, (14, 12), cdt)
    lyot_r = rng.random((14, 12)).astype(rdt)
    lyot_c = crand(rng, (14, 12), cdt)
    for kind in ('mdft', 'czt'):
        ex_r = propagation.prepare_executor(0.25, (14, 12), 0.8, (10, 16), 0.55, 20.0, kind=kind)
        ex_o = O.prepare_executor(0.25, (14, 12), 0.8, (10, 16), 0.55, 20.0, kind=kind, rdtype=rdt)
        for fname, fpm in (('real', rng.random((10, 16)).astype(rdt)), ('cplx', crand(rng, (10, 16), cdt))):
            t2 = tag + f'{kind} fpm={fname} '
            got = O.to_fpm_and_back(w, fpm, ex_o, True)
            ref = propagation.to_fpm_and_back(w, fpm, ex_r, return_more=True)
            for nm, u, v in zip(('next', 'at_fpm', 'after_fpm'), got, ref):
                check(t2 + 'to_fpm_and_back ' + nm, u, v, eps)
            got = O.to_fpm_and_back_adjoint(g, fpm, ex_o, True, ref[1])
            refa = propagation.to_fpm_and_back_adjoint(g, fpm, ex_r, return_more=True, return_fpm_grad=True, field_at_fpm=ref[1])
            for nm, u, v in zip(('Eabar', 'Ebbar', 'inter', 'fpm_bar'), got, refa):
                check(t2 + 'to_fpm_and_back_adjoint ' + nm, u, v, eps)
            for lname, lyot in (('none', None), ('real', lyot_r), ('cplx', lyot_c)):
                got = O.babinet(w, lyot, fpm, ex_o, True)
                refb = propagation.babinet(w, lyot, fpm, ex_r, return_more=True)
                for nm, u, v in zip(('after_lyot', 'at_fpm', 'after_fpm', 'at_lyot'), got, refb):
                    check(t2 + f'babinet lyot={lname} ' + nm, u, v, eps)
                got = O.babinet_adjoint(g, lyot, fpm, ex_o, refb[1], refb[3])
                refg = propagation.babinet_adjoint(g, lyot, fpm, ex_r, field_at_fpm=refb[1], field_at_lyot=refb[3],
                                                   return_fpm_grad=True, return_lyot_grad=True)
                for nm, u, v in zip(('abar', 'fpm_bar', 'lyot_bar'), got, refg):
                    check(t2 + f'babinet_adjoint lyot={lname} ' + nm, u, v, eps)
        # multi-resolution stack, vortex mask
        mr = propagation.prepare_multiresolution(0.1, (14, 12), 2.0, (12, 18), 0.55, 10.0, num_levels=3, fine_samples=10, kind=kind)
        mo = O.prepare_multiresolution(0.1, (14, 12), 2.0, (12, 18), 0.55, 10.0, num_levels=3, fine_samples=10, kind=kind, rdtype=rdt)
        for k in range(3):
            check(tag + f'{kind} multires window[{k}]', mo.windows[k], mr.windows[k], eps)
            check(tag + f'{kind} multires xf[{k}]', mo.xf[k], mr.xf[k], eps)
            check(tag + f'{kind} multires yf[{k}]', mo.yf[k], mr.yf[k], eps)
        vo, vr = O.vortex_phase_mask(2), propagation.vortex_phase_mask(2)
        check(tag + 'vortex mask', vo(mo.xf[1], mo.yf[1]), vr(mr.xf[1], mr.yf[1]), eps)
        got = O.to_fpm_and_back_multiresolution(w, vo, mo, True)
        ref = propagation.to_fpm_and_back_multiresolution(w, vr, mr, return_more=True)
        check(tag + f'{kind} multires forward', got[0], ref[0], eps)
        for k in range(3):
            check(tag + f'{kind} multires at_fpm[{k}]', got[1][k], ref[1][k], eps)
            check(tag + f'{kind} multires after_fpm[{k}]', got[2][k], ref[2][k], eps)
        got = O.to_fpm_and_back_multiresolution_adjoint(g, vo, mo, ref[1])
        refa = propagation.to_fpm_and_back_multiresolution_adjoint(g, vr, mr, return_more=True, return_fpm_grad=True, field_at_fpm=ref[1])
        check(tag + f'{kind} multires adjoint', got[0], refa[0], eps)
        for k in range(3):
            check(tag + f'{kind} multires Ebbar[{k}]', got[1][k], refa[1][k], eps)
            check(tag + f'{kind} multires inter[{k}]', got[2][k], refa[2][k], eps)
            check(tag + f'{kind} multires fpm_bar[{k}]', got[3][k], refa[3][k], eps)

    # ---- pupil synthesis by recurrence (rank 3)
    from prysm import coordinates as pcoord, geometry as pgeom
    from prysm.polynomials import jacobi as pjacobi, jacobi_seq as pjacobi_seq, zernike_nm as pz_nm, zernike_sum as pz_sum
    for shp, kw in (((9, 12), dict(dx=0.25)), (16, dict(diameter=2.0))):
        xr, yr = pcoord.make_xy_grid(shp, **kw)
        xo, yo = O.make_xy_grid(shp, dtype=rdt, **kw)
        check(tag + f'make_xy_grid {shp} x', xo, xr)
        check(tag + f'make_xy_grid {shp} y', yo, yr)
    rr, tr = pcoord.cart_to_polar(xr, yr)
    ro, to = O.cart_to_polar(xo, yo)
    check(tag + 'cart_to_polar r', ro, rr)
    check(tag + 'cart_to_polar t', to, tr)
    check(tag + 'circle', O.circle(0.7, ro), pgeom.circle(0.7, rr))
    check(tag + 'antialias', O.antialias(ro - 0.7, 2.0 / 16), pgeom.antialias(rr - rdt(0.7), 2.0 / 16), eps)
    xs = np.linspace(-1, 1, 41).astype(rdt)
    for (al, be) in ((0, 0), (0, 3), (1.5, 0.5), (-0.5, -0.5)):
        check(tag + f'jacobi_seq a={al} b={be}', O.jacobi_seq([0, 1, 2, 5, 9], al, be, xs), pjacobi_seq([0, 1, 2, 5, 9], al, be, xs), eps)
        check(tag + f'jacobi n=7 a={al} b={be}', O.jacobi(7, al, be, xs), pjacobi(7, al, be, xs), eps)
    nms = [noll_to_nm(j) for j in range(1, 38)]
    rn = ro / ro.max()
    for nrm in (True, False):
        check(tag + f'zernike_nm_seq norm={nrm}', O.zernike_nm_seq(nms, rn, to, nrm), zernike_nm_seq(nms, rn, tr, norm=nrm), eps * 10)
    check(tag + 'zernike_nm (single) == seq', O.zernike_nm_seq([(4, -2)], rn, to)[0], pz_nm(4, -2, rn, tr), eps * 10)
    cz = rng.standard_normal(37).astype(rdt)
    cz[3] = 0
    check(tag + 'zernike_sum', O.zernike_sum(cz, nms, xo, yo), pz_sum(cz, nms, xr, yr), eps * 10)

    # ---- detector sampling
    from prysm import detector as pdet
    img = rng.random((12, 18)).astype(rdt)
    for fac in (2, 3, (2, 3), (4, 6)):
        for mode in ('avg', 'sum'):
            check(tag + f'bindown {fac} {mode}', O.bindown(img, fac, mode), pdet.bindown(img, fac, mode), eps)
            check(tag + f'tile {fac} {mode}', O.tile(img, fac, mode), pdet.tile(img, fac, mode), eps)
    gfx, gfy = O.transfer_function_grids((12, 18), 2.0, False, rdt)[:2]
    check(tag + 'pixel_ft', O.pixel_ft(gfx, gfy, 3.0, 2.5), pdet.pixel_ft(gfx, gfy, 3.0, 2.5), eps)
    check(tag + 'olpf_ft', O.olpf_ft(gfx, gfy, 0.7, 0.9), pdet.olpf_ft(gfx, gfy, 0.7, 0.9), eps)
    # ---- measured focal-plane mask resampling
    mm = crand(rng, (21, 17), cdt)
    qx = (rng.random((9, 11)) * 12 - 6).astype(rdt)
    qy = (rng.random((9, 11)) * 12 - 6).astype(rdt)
    for kw in (dict(charge=2), dict(fill=0.25), dict()):
        check(tag + f'prepare_measured_fpm {kw}', O.prepare_measured_fpm(mm, 0.4, (0.3, -0.2), **kw)(qx, qy),
              propagation.prepare_measured_fpm(mm, 0.4, center=(0.3, -0.2), **kw)(qx, qy), eps)
    # ---- image-chain consumers (rank 4)
    from prysm import convolution as pconv
    ob = rng.random((12, 10)).astype(rdt)
    ps = rng.random((12, 10)).astype(rdt)
    obc = crand(rng, (12, 10), cdt)
    check(tag + 'conv real', O.conv(ob, ps), pconv.conv(ob, ps), eps)
    check(tag + 'conv complex object', O.conv(obc, ps), pconv.conv(obc, ps), eps)
    tf1 = rng.random((12, 10)).astype(rdt)
    tf2 = crand(rng, (12, 10), cdt)
    for sh in (False, True):
        check(tag + f'apply_transfer_functions arrays shift={sh}', O.apply_transfer_functions(ob, [tf1, tf2], sh),
              pconv.apply_transfer_functions(ob, 0.5, [tf1, tf2], shift=sh), eps)
        seen = {}

        def probe(fx, fy, fr, ft):
            seen.update(fx=fx, fy=fy, fr=fr, ft=ft)
            return np.ones(ob.shape, dtype=rdt)
        pconv.apply_transfer_functions(ob, 0.5, [probe], shift=sh)
        for nm, v in zip(('fx', 'fy', 'fr', 'ft'), O.transfer_function_grids(ob.shape, 0.5, sh, rdt)):
            check(tag + f'tf grid {nm} shift={sh}', v, seen[nm], eps)
    for zoom in (0.5, 2, (2, 1.5)):
        check(tag + f'fourier_resample zoom={zoom}', O.fourier_resample(ob, zoom, rdt), fttools.fourier_resample(ob, zoom), eps * 2)
    check(tag + 'fourier_resample complex', O.fourier_resample(obc, 2, rdt), fttools.fourier_resample(obc, 2), eps * 2)

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
