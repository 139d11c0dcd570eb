"""CPU: the oracle restatement must reproduce the reference's own outputs (tests/golden/*.npz,
written by oracle/make_golden.py from the unmodified reference).  This is what pins the oracle
on boxes where /root/reference does not exist."""
import numpy as np
import pytest

import prysm_oracle as O
from conftest import rel_linf, load_golden

HeNe = 0.6328
FOCUS_CASES = list(range(9))


@pytest.mark.parametrize('i', FOCUS_CASES)
def test_focus_family(small, i):
    a, Q = small[f'focus{i}_in'], float(small[f'focus{i}_Q'])
    Q = int(Q) if Q == int(Q) else Q
    g = small[f'focus{i}_gin']
    assert rel_linf(O.focus(a, Q), small[f'focus{i}_focus']) < 1e-14
    assert rel_linf(O.unfocus(a, Q), small[f'focus{i}_unfocus']) < 1e-14
    assert rel_linf(O.focus_adjoint(g, Q), small[f'focus{i}_focus_adjoint']) < 1e-14
    assert rel_linf(O.unfocus_adjoint(g, Q), small[f'focus{i}_unfocus_adjoint']) < 1e-14


def test_wavefront_chain(small):
    amp, opd, dx = small['wf_amp'], small['wf_opd'], float(small['wf_dx'])
    amp_o, opd_o, dx_o = O.synthetic_pupil(64, np.float64)
    assert (amp_o == amp).all() and dx_o == dx
    assert np.abs((opd_o - opd) * amp).max() < 1e-9
    field = O.from_amp_and_phase(amp, opd, HeNe)
    assert rel_linf(field, small['wf_field']) < 1e-14
    psf = O.focus(field, 2)
    assert rel_linf(psf, small['wf_psf_field']) < 1e-13
    assert O.pupil_sample_to_psf_sample(dx, psf.shape[1], HeNe, 100.0) == pytest.approx(float(small['wf_psf_dx']), rel=1e-15)
    I = O.intensity(psf)
    assert rel_linf(I, small['wf_psf_intensity']) < 1e-13
    assert rel_linf(O.unfocus(psf, 1), small['wf_back_field']) < 1e-13
    pdx = float(small['wf_psf_dx'])
    mtf, df = O.mtf_from_psf(I, pdx)
    assert rel_linf(mtf, small['wf_mtf']) < 1e-12 and df == pytest.approx(float(small['wf_mtf_df']))
    assert rel_linf(O.otf_from_psf(I, pdx)[0], small['wf_otf']) < 1e-12
    assert np.allclose(O.centroid(I, pdx), small['wf_centroid'], rtol=1e-10, atol=1e-12)
    assert np.allclose(O.centroid(I, unit='pixels'), small['wf_centroid_px'], rtol=1e-12)
    g = O.fftrange(64) * dx
    x, y = np.meshgrid(g, g)
    assert rel_linf(O.thin_lens(250.0, HeNe, x, y), small['lens']) < 1e-14


def test_angular_spectrum(small):
    f = small['as_in']
    for Q in (1, 2):
        assert rel_linf(O.angular_spectrum(f, HeNe, 0.05, 12.5, Q), small[f'as_Q{Q}']) < 1e-14
    assert rel_linf(O.angular_spectrum_transfer_function((24, 32), HeNe, 0.05, 12.5), small['as_tf']) < 1e-15
    assert rel_linf(O.angular_spectrum(f, HeNe, 0.05, 12.5, tf=small['as_tf']), small['as_with_tf']) < 1e-14
    assert rel_linf(O.angular_spectrum_adjoint(small['as_gin'], HeNe, 0.05, 12.5, 2), small['as_adjoint_Q2']) < 1e-14
    assert rel_linf(O.angular_spectrum(small['as9_in'], HeNe, 0.05, 3.0, 1), small['as9_Q1']) < 1e-14
    assert rel_linf(O.angular_spectrum(small['as9_in'], HeNe, 0.05, 3.0, 1.5), small['as9_Q15']) < 1e-14


@pytest.mark.parametrize('i', range(4))
@pytest.mark.parametrize('kind', ['mdft', 'czt'])
def test_executors(small, i, kind):
    a, g = small[f'ex{i}_in'], small[f'ex{i}_gin']
    pdx, fdx, wvl, efl, sx, sy = small[f'ex{i}_params']
    ex = O.prepare_executor(pdx, a.shape, fdx, g.shape, wvl, efl, (sx, sy), kind)
    assert rel_linf(ex(a), small[f'ex{i}_{kind}_fwd']) < 1e-13
    assert rel_linf(ex.adjoint(g), small[f'ex{i}_{kind}_adj']) < 1e-13


@pytest.mark.parametrize('i', range(2))
def test_fftdft(small, i):
    a, g = small[f'fd{i}_in'], small[f'fd{i}_gin']
    pdx, fdx, wvl, efl, sx, sy = small[f'fd{i}_params']
    ex = O.prepare_executor(pdx, a.shape, fdx, g.shape, wvl, efl, (sx, sy), 'fftdft')
    assert rel_linf(ex(a), small[f'fd{i}_fwd']) < 1e-13
    assert rel_linf(ex.adjoint(g), small[f'fd{i}_adj']) < 1e-13


def test_executor_equivalences():
    """The identities the reference's tests pin (tests/test_fttools.py:140-182, test_propagation.py:98-117)."""
    rng = np.random.default_rng(3)
    a = rng.standard_normal((16, 16)) + 1j * rng.standard_normal((16, 16))
    K = 32
    fdx = HeNe * 100.0 / (0.1 * K)
    m = O.prepare_executor(0.1, 16, fdx, 32, HeNe, 100.0, kind='mdft')(a)
    c = O.prepare_executor(0.1, 16, fdx, 32, HeNe, 100.0, kind='czt')(a)
    f = O.prepare_executor(0.1, 16, fdx, 32, HeNe, 100.0, kind='fftdft')(a)
    assert rel_linf(c, m) < 1e-12 and rel_linf(f, m) < 1e-12
    assert rel_linf(m, O.focus(a, 2)) < 1e-12  # FFT-equivalent sampling


def test_mode_sum(small):
    assert rel_linf(O.sum_of_2d_modes(small['modes'], small['weights']), small['modes_sum']) < 1e-15


def test_c1_full_reference_window():
    """BASELINE config C1 (256^2 -> 512^2, fp64) against the reference's stored window."""
    g = load_golden('full_c1.npz')
    amp, opd, dx = O.synthetic_pupil(256, np.float64)
    psf = O.focus(O.from_amp_and_phase(amp, opd, HeNe), 2)
    cy = psf.shape[0] // 2
    assert rel_linf(psf[cy - 32:cy + 32, cy - 32:cy + 32], g['field_win']) < 1e-9
    I = O.intensity(psf)
    assert I.sum() == pytest.approx(float(g['I_sum']), rel=1e-10)
    assert I.sum() == pytest.approx(float(g['E_in']), rel=1e-12)  # ortho FFT conserves energy


def test_c4_full_reference_samples():
    """BASELINE config C4 (2048^2 -> 2048^2 CZT per wavelength, weighted incoherent sum) against the reference's stored
    fp64 samples at two wavelengths."""
    g = load_golden('full_c4.npz')
    amp, opd, dx = O.synthetic_pupil(2048, np.float32)     # the golden inputs: OPD rounded to float32 once
    opd = opd.astype(np.float64)
    tot = 0
    for w, wt in ((0.5, 0.25), (0.7, 0.75)):
        ex = O.prepare_executor(dx, (2048, 2048), 2.5, (2048, 2048), w, 100.0, kind='czt')
        f = ex(O.from_amp_and_phase(amp, opd, w))
        tag = f'w{int(w * 10)}_'
        assert np.abs(f[::64, ::64] - g[tag + 'field_stride']).max() / float(g[tag + 'absmax']) < 1e-9
        cy = 1024
        assert np.abs(f[cy - 32:cy + 32, cy - 32:cy + 32] - g[tag + 'field_win']).max() / float(g[tag + 'absmax']) < 1e-9
        I = O.intensity(f)
        assert I.sum() == pytest.approx(float(g[tag + 'I_sum']), rel=1e-9)
        tot = tot + wt * I
    assert np.abs(tot[::64, ::64] - g['sum_stride']).max() / float(g['sum_max']) < 1e-9
    assert tot.sum() == pytest.approx(float(g['sum_total']), rel=1e-9)


def test_c5_full_reference_samples():
    """BASELINE config C5: one 4096^2 plane -- screen multiply then free_space(dz = 5 mm) -- against the reference."""
    g = load_golden('full_c5.npz')
    amp, opd, dx = O.synthetic_pupil(4096, np.float32)     # the golden inputs: OPD and screen rounded to float32 once
    field = O.from_amp_and_phase(amp, opd.astype(np.float64), HeNe)
    phi32 = np.random.default_rng(1000).normal(0, 0.1, (4096, 4096)).astype(np.float32)
    scr_opd = (phi32.astype(np.float64) * (HeNe * 1e3 / (2 * np.pi))).astype(np.float32).astype(np.float64)
    out = O.angular_spectrum(field * O.phase_screen(scr_opd, HeNe), HeNe, dx, 5.0, 1)
    den = float(g['absmax'])
    assert np.abs(out[::128, ::128] - g['field_stride']).max() / den < 1e-9
    assert np.abs(out[2048 - 32:2048 + 32, 2048 - 32:2048 + 32] - g['field_win']).max() / den < 1e-9
    assert np.abs(out[2048, 1000:1100] - g['edge']).max() / den < 1e-9
    assert (np.abs(out) ** 2).sum() == pytest.approx(float(g['E_out']), rel=1e-10)
    assert float(g['E_out']) == pytest.approx(float(g['E_in']), rel=1e-10)        # |TF| = 1, |screen| = 1
    foc = O.prepare_executor(dx, (4096, 4096), float(g['focal_dx']), (512, 512), HeNe, 100.0, kind='czt')(out)   # final CZT focus
    fden = float(g['focus_absmax'])
    assert np.abs(foc[256 - 32:256 + 32, 256 - 32:256 + 32] - g['focus_win']).max() / fden < 1e-9
    assert np.abs(foc[::16, ::16] - g['focus_stride']).max() / fden < 1e-9
