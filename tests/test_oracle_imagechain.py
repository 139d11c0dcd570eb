"""CPU: the oracle's image-chain restatement (conv, apply_transfer_functions, fourier_resample)
against the reference's outputs in tests/golden/imagechain.npz (oracle/make_golden.py imagechain)."""
import numpy as np
import pytest

import prysm_oracle as O
from conftest import rel_linf, load_golden

ZOOMS = (0.5, 2, (2, 1.5))


@pytest.fixture(scope='module')
def gold():
    return load_golden('imagechain.npz')


def test_conv(gold):
    g = gold
    assert rel_linf(O.conv(g['obj'], g['psf']), g['conv_real']) < 1e-14
    assert rel_linf(O.conv(g['obj_c'], g['psf']), g['conv_cplx']) < 1e-14
    assert rel_linf(O.conv(g['odd_obj'], g['odd_psf']), g['conv_odd']) < 1e-14
    # reference tests/test_convolution.py:10-17: a centred delta is the identity
    obj = np.arange(25, dtype=float).reshape(5, 5)
    d = np.zeros_like(obj)
    d[2, 2] = 1
    assert np.allclose(O.conv(obj, d), obj, atol=1e-12)


@pytest.mark.parametrize('shift', [False, True])
def test_apply_transfer_functions(gold, shift):
    g = gold
    s = int(shift)
    assert rel_linf(O.apply_transfer_functions(g['obj'], [g['tf1'], g['tf2']], shift), g[f'atf_shift{s}']) < 1e-14
    assert rel_linf(O.apply_transfer_functions(g['obj_c'], [g['tf1']], shift), g[f'atf_c_shift{s}']) < 1e-14
    fx, fy, fr, ft = O.transfer_function_grids(g['obj'].shape, 0.5, shift)
    for nm, v in (('fx', fx), ('fy', fy), ('fr', fr), ('ft', ft)):
        assert v.shape == g[f'grid_{nm}_shift{s}'].shape and rel_linf(v, g[f'grid_{nm}_shift{s}']) < 1e-15
    assert rel_linf(O.apply_transfer_functions(g['obj'], [np.exp(-(fr / 0.7) ** 2)], shift), g[f'atf_callable_shift{s}']) < 1e-14


def test_fourier_resample(gold):
    g = gold
    for i, z in enumerate(ZOOMS):
        out = O.fourier_resample(g['obj'], z)
        assert out.shape == g[f'resample{i}'].shape and rel_linf(out, g[f'resample{i}']) < 1e-13
    assert rel_linf(O.fourier_resample(g['obj_c'], 2), g['resample_c']) < 1e-13
    a = g['obj']
    assert O.fourier_resample(a, 1) is a
    with pytest.raises(ValueError):
        O.fourier_resample(g['obj'], -1)
    assert np.allclose(O.fourier_resample(np.ones((8, 8)), (2, 3)), 1, atol=1e-12)      # tests/test_fttools.py:229-235


def test_detector_sampling(gold):
    g = gold
    for i, fac in enumerate((2, 3, (2, 3), (4, 6))):
        assert rel_linf(O.bindown(g['obj'], fac, 'avg'), g[f'bin{i}_avg']) < 1e-15
        assert rel_linf(O.bindown(g['obj'], fac, 'sum'), g[f'bin{i}_sum']) < 1e-15
        assert rel_linf(O.tile(g['psf'][:6, :5], fac, 'sum'), g[f'tile{i}_sum']) < 1e-15
    fx, fy = O.transfer_function_grids(g['obj'].shape, 2.0, False)[:2]
    assert rel_linf(O.pixel_ft(fx, fy, 3.0, 2.5), g['pixel_ft']) < 1e-15
    assert rel_linf(O.olpf_ft(fx, fy, 0.7, 0.9), g['olpf_ft']) < 1e-15
    a = np.arange(24.0).reshape(4, 6)
    assert np.vdot(O.bindown(a, 2, 'sum'), np.ones((2, 3))) == np.vdot(a, O.tile(np.ones((2, 3)), 2, 'avg'))   # adjoint pair
    with pytest.raises(ValueError):
        O.bindown(a, 2, 'median')
