"""CPU: the oracle's adjoint twins and Lyot-coronagraph compositions must reproduce the
reference's own outputs stored in tests/golden/coronagraph.npz (oracle/make_golden.py coronagraph)."""
import numpy as np
import pytest

import prysm_oracle as O
from conftest import rel_linf, load_golden

HeNe = 0.6328


@pytest.fixture(scope='module')
def gold():
    return load_golden('coronagraph.npz')


def test_elementwise_adjoints(gold):
    g = gold
    field = O.from_amp_and_phase(g['ea_amp'], g['ea_opd'], 0.55)
    assert rel_linf(O.intensity_adjoint(field, g['ea_ibar']), g['ea_intensity_adjoint']) < 1e-14
    assert rel_linf(O.from_amp_and_phase_adjoint_phase(field, g['ea_bar'], 0.55), g['ea_phase']) < 1e-14
    assert rel_linf(O.from_amp_and_phase_adjoint_amp(field, g['ea_bar'], 0.55), g['ea_amp_nophase']) < 1e-13
    assert rel_linf(O.from_amp_and_phase_adjoint_amp(field, g['ea_bar'], 0.55, g['ea_opd']), g['ea_amp_phase']) < 1e-14
    assert O.from_amp_and_phase_adjoint_amp(field, g['ea_bar'], 0.55)[2, 3] == 0      # zero amplitude -> zero gradient
    x, y = np.meshgrid(O.fftrange(24) * 0.1, O.fftrange(18) * 0.1)
    assert O.thin_lens_adjoint(250.0, 0.55, x, y, g['ea_bar']) == pytest.approx(float(g['ea_lens_adjoint']), rel=1e-12)
    assert rel_linf(O.sum_of_2d_modes_adjoint(g['ma_modes'], g['ea_ibar']), g['ma_out']) < 1e-14


def test_otf_adjoints(gold):
    g = gold
    D = O.transform_psf(g['oa_psf'], 1.5)[0]
    assert rel_linf(O.mtf_from_psf_adjoint(g['oa_rbar'], D), g['oa_mtf']) < 1e-13
    assert rel_linf(O.ptf_from_psf_adjoint(g['oa_rbar'], D), g['oa_ptf']) < 1e-13
    assert rel_linf(O.otf_from_psf_adjoint(g['oa_cbar'], D), g['oa_otf']) < 1e-13
    assert rel_linf(O.encircled_energy_adjoint([0.3, -1.2], D, 1.5, [2.0, 7.5]), g['oa_ee']) < 1e-12


@pytest.mark.parametrize('kind', ['mdft', 'czt'])
@pytest.mark.parametrize('fname', ['real', 'cplx'])
def test_single_executor_compositions(gold, kind, fname):
    g = gold
    pdx, fdx, wvl, efl = g['co_params']
    w, gb, lyot, fpm = g['co_w'], g['co_g'], g['co_lyot'], g[f'co_fpm_{fname}']
    ex = O.prepare_executor(pdx, w.shape, fdx, fpm.shape, wvl, efl, kind=kind)
    t = f'co_{kind}_{fname}_'
    nxt, at, after = O.to_fpm_and_back(w, fpm, ex, True)
    assert rel_linf(nxt, g[t + 'next']) < 1e-12 and rel_linf(at, g[t + 'at_fpm']) < 1e-12
    assert rel_linf(after, g[t + 'after_fpm']) < 1e-12
    Ea, Eb, it, fb = O.to_fpm_and_back_adjoint(gb, fpm, ex, True, at)
    for u, nm in ((Ea, 'Eabar'), (Eb, 'Ebbar'), (it, 'inter'), (fb, 'fpm_bar')):
        assert rel_linf(u, g[t + nm]) < 1e-12, nm
    assert np.iscomplexobj(fb) == (fname == 'cplx')
    al, at2, _, atl = O.babinet(w, lyot, fpm, ex, True)
    assert rel_linf(al, g[t + 'bab_after_lyot']) < 1e-12 and rel_linf(atl, g[t + 'bab_at_lyot']) < 1e-12
    ab, fbb, lb = O.babinet_adjoint(gb, lyot, fpm, ex, at2, atl)
    for u, nm in ((ab, 'bab_abar'), (fbb, 'bab_fpm_bar'), (lb, 'bab_lyot_bar')):
        assert rel_linf(u, g[t + nm]) < 1e-12, nm
    with pytest.raises(ValueError):
        O.to_fpm_and_back_adjoint(gb, fpm, ex, True, None)


@pytest.mark.parametrize('kind', ['mdft', 'czt'])
def test_multiresolution(gold, kind):
    g = gold
    pdx, fdx, wvl, efl = g['mr_params']
    x, y = g['mr_x'], g['mr_y']
    mex = O.prepare_multiresolution(pdx, 64, fdx, 32, wvl, efl, num_levels=3, fine_samples=32, kind=kind)
    assert len(mex) == 3
    for k in range(3):
        assert rel_linf(mex.windows[k], g[f'mr_win{k}']) < 1e-14
        assert rel_linf(mex.xf[k], g[f'mr_xf{k}']) < 1e-15 and rel_linf(mex.yf[k], g[f'mr_yf{k}']) < 1e-15
    fpm = O.vortex_phase_mask(2)
    assert rel_linf(fpm(mex.xf[1], mex.yf[1]), g['mr_vortex1']) < 1e-15
    t = f'mr_{kind}_'
    out, at, after = O.to_fpm_and_back_multiresolution(x, fpm, mex, True)
    assert rel_linf(out, g[t + 'out']) < 1e-12
    Ea, Ebs, its, fbs = O.to_fpm_and_back_multiresolution_adjoint(y, fpm, mex, at)
    assert rel_linf(Ea, g[t + 'Eabar']) < 1e-12
    for k in range(3):
        assert rel_linf(at[k], g[t + f'at{k}']) < 1e-12 and rel_linf(after[k], g[t + f'after{k}']) < 1e-12
        assert rel_linf(Ebs[k], g[t + f'Ebbar{k}']) < 1e-12 and rel_linf(its[k], g[t + f'inter{k}']) < 1e-12
        assert rel_linf(fbs[k], g[t + f'fpm_bar{k}']) < 1e-12
    # the reference's own identity: <A x, y> == <x, A^H y>  (tests/test_propagation.py:544-556)
    assert np.vdot(out, y) == pytest.approx(np.vdot(x, Ea), rel=1e-10)


def test_windows_partition_unity():
    """Where two levels overlap on a common point set the windows telescope to one (dft.py:283-292): check on the
    coarsest grid by evaluating every level's window function there."""
    mex = O.prepare_multiresolution(0.1, 32, 2.0, 48, HeNe, 10.0, num_levels=4, fine_samples=24)
    r = np.hypot(mex.xf[0], mex.yf[0])
    halves = [min(x.shape) / 2.0 * (2.0 / 4.0 ** k) for k, x in enumerate(mex.xf)]
    tot = 0
    for k in range(4):
        here = 1.0 if k == 0 else O.cumulative_window(r, 0.2 * halves[k], 0.7 * halves[k])
        nxt = 0.0 if k == 3 else O.cumulative_window(r, 0.2 * halves[k + 1], 0.7 * halves[k + 1])
        tot = tot + (here - nxt)
    assert np.allclose(tot, 1.0, atol=1e-15)
    with pytest.raises(TypeError):
        O.vortex_phase_mask(2.5)
    O.vortex_phase_mask(np.int64(2))


def test_measured_fpm_resampling(gold):
    """prepare_measured_fpm (coronagraph.py:135-209) incl. the reference's own identities (tests/test_propagation.py:648-676)."""
    g = gold
    mm, qx, qy = g['mf_map'], g['mf_qx'], g['mf_qy']
    assert rel_linf(O.prepare_measured_fpm(mm, 0.4, (0.3, -0.2), charge=2)(qx, qy), g['mf_vortex']) < 1e-14
    assert rel_linf(O.prepare_measured_fpm(mm, 0.4, (0.3, -0.2), fill=0.25)(qx, qy), g['mf_scalar']) < 1e-14
    assert rel_linf(O.prepare_measured_fpm(mm, 0.4)(qx, qy), g['mf_default']) < 1e-14
    x, y = O.make_xy_grid(129, dx=0.4)
    meas = np.exp(1j * 2 * np.arctan2(y, x))
    f = O.prepare_measured_fpm(meas, 0.4, charge=2)
    assert np.allclose(f(x, y), meas, atol=1e-12)
    far = np.full((1, 1), 1e5)
    assert np.allclose(f(far, far), np.exp(1j * 2 * np.arctan2(far, far)), atol=1e-12)
