"""CPU: size-independent properties of the oracle itself, on random shapes (hypothesis) -- the identities the
reference's own suite uses to pin the path: adjoint pairs, round trips, executor equivalences,
energy conservation, linearity.  These are the same properties the GPU suite checks at sizes the oracle cannot reach."""
import numpy as np
import pytest
from hypothesis import given, settings, strategies as st, HealthCheck

import prysm_oracle as O
from conftest import rel_linf

HeNe = 0.6328
SETTINGS = dict(max_examples=25, deadline=None, suppress_health_check=[HealthCheck.too_slow])
shapes = st.tuples(st.integers(2, 24), st.integers(2, 24))
Qs = st.sampled_from([1, 1.5, 2, 3])


def crand(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@given(shape=shapes, Q=Qs, seed=st.integers(0, 2 ** 16))
@settings(**SETTINGS)
def test_focus_family_adjoints_and_round_trip(shape, Q, seed):
    rng = np.random.default_rng(seed)
    x = crand(rng, shape)
    fx = O.focus(x, Q)
    y = crand(rng, fx.shape)
    assert np.vdot(fx, y) == pytest.approx(np.vdot(x, O.focus_adjoint(y, Q)), rel=1e-10, abs=1e-10)
    assert np.vdot(O.unfocus(x, Q), y) == pytest.approx(np.vdot(x, O.unfocus_adjoint(y, Q)), rel=1e-10, abs=1e-10)
    assert rel_linf(O.unfocus(O.focus(x, 1), 1), x) < 1e-12                          # tests/test_propagation.py:24-29
    assert np.sum(np.abs(fx) ** 2) == pytest.approx(np.sum(np.abs(x) ** 2), rel=1e-12)   # ortho norm conserves energy
    assert rel_linf(O.crop_center(O.pad2d(x, Q), x.shape), x) == 0                    # tests/test_fttools.py:88-93


@given(shape=shapes, Q=st.sampled_from([1, 2]), z=st.floats(-20, 20), seed=st.integers(0, 2 ** 16))
@settings(**SETTINGS)
def test_angular_spectrum_adjoint_and_unitarity(shape, Q, z, seed):
    rng = np.random.default_rng(seed)
    x = crand(rng, shape)
    fwd = O.angular_spectrum(x, HeNe, 0.05, z, Q)
    y = crand(rng, fwd.shape)
    assert np.vdot(fwd, y) == pytest.approx(np.vdot(x, O.angular_spectrum_adjoint(y, HeNe, 0.05, z, Q)), rel=1e-10, abs=1e-10)
    if Q == 1:
        assert rel_linf(O.angular_spectrum(fwd, HeNe, 0.05, -z, 1), x) < 1e-11         # |TF| = 1
        assert rel_linf(O.angular_spectrum(x, HeNe, 0.05, 0.0, 1), x) < 1e-13          # tests/test_propagation.py:210-218


@given(n=st.integers(4, 20), m=st.integers(4, 20), kind=st.sampled_from(['mdft', 'czt']),
       shift=st.tuples(st.floats(-3, 3), st.floats(-3, 3)), seed=st.integers(0, 2 ** 16))
@settings(**SETTINGS)
def test_executor_adjoint_and_czt_equals_mdft(n, m, kind, shift, seed):
    rng = np.random.default_rng(seed)
    a = crand(rng, (n, n + 1))
    g = crand(rng, (m + 1, m))
    args = (0.1, a.shape, 1.7, g.shape, HeNe, 50.0, shift)
    ex = O.prepare_executor(*args, kind)
    assert np.vdot(ex(a), g) == pytest.approx(np.vdot(a, ex.adjoint(g)), rel=1e-10, abs=1e-10)
    other = O.prepare_executor(*args, 'czt' if kind == 'mdft' else 'mdft')
    assert rel_linf(ex(a), other(a)) < 1e-10                                          # tests/test_fttools.py:140-155
    assert rel_linf(ex.adjoint(g), other.adjoint(g)) < 1e-10


@given(n=st.sampled_from([8, 12, 16]), K=st.sampled_from([16, 24, 32]), seed=st.integers(0, 2 ** 16))
@settings(**SETTINGS)
def test_fft_equivalent_sampling_all_executors_agree(n, K, seed):
    """focal_dx = wvl*efl/(pupil_dx*K): MDFT == CZT == FFTDFT == focus cropped (tests/test_propagation.py:98-117)."""
    rng = np.random.default_rng(seed)
    a = crand(rng, (n, n))
    fdx = HeNe * 100.0 / (0.1 * K)
    outs = [O.prepare_executor(0.1, n, fdx, K, HeNe, 100.0, kind=k)(a) for k in ('mdft', 'czt', 'fftdft')]
    assert rel_linf(outs[1], outs[0]) < 1e-11 and rel_linf(outs[2], outs[0]) < 1e-11
    ref = O.focus(O.pad2d(a, out_shape=(K, K)), 1) if K != n else O.focus(a, 1)
    assert rel_linf(outs[0], ref) < 1e-11


@given(shape=shapes, fshape=shapes, seed=st.integers(0, 2 ** 16), cplx=st.booleans())
@settings(**SETTINGS)
def test_coronagraph_compositions_are_adjoint_pairs(shape, fshape, seed, cplx):
    rng = np.random.default_rng(seed)
    x, y = crand(rng, shape), crand(rng, shape)
    fpm = crand(rng, fshape) if cplx else rng.standard_normal(fshape)
    lyot = rng.standard_normal(shape)
    ex = O.prepare_executor(0.25, shape, 0.1, fshape, HeNe, 10.0)
    lhs = np.vdot(O.to_fpm_and_back(x, fpm, ex), y)
    assert lhs == pytest.approx(np.vdot(x, O.to_fpm_and_back_adjoint(y, fpm, ex)[0]), rel=1e-10, abs=1e-10)
    lhs = np.vdot(O.babinet(x, lyot, fpm, ex), y)
    assert lhs == pytest.approx(np.vdot(x, O.babinet_adjoint(y, lyot, fpm, ex)[0]), rel=1e-10, abs=1e-10)
    # Babinet: lyot * (x - c(1 - fpm))
    assert rel_linf(O.babinet(x, lyot, fpm, ex), lyot * (x - O.to_fpm_and_back(x, 1 - fpm, ex))) < 1e-13


@given(m=st.integers(1, 6), n=st.integers(1, 6), fy=st.integers(1, 4), fx=st.integers(1, 4), seed=st.integers(0, 2 ** 16))
@settings(**SETTINGS)
def test_image_chain_properties(m, n, fy, fx, seed):
    rng = np.random.default_rng(seed)
    a = rng.random((m * fy, n * fx))
    b = rng.random((m, n))
    assert np.vdot(O.bindown(a, (fy, fx), 'sum'), b) == pytest.approx(np.vdot(a, O.tile(b, (fy, fx), 'avg')), rel=1e-12)
    assert O.bindown(a, (fy, fx), 'avg').sum() * fy * fx == pytest.approx(a.sum(), rel=1e-12)
    o, h1, h2 = rng.random(a.shape), rng.random(a.shape), rng.random(a.shape)
    assert rel_linf(O.conv(o, h1 + 2 * h2), O.conv(o, h1) + 2 * O.conv(o, h2)) < 1e-12     # linear in the PSF
    assert rel_linf(O.conv(o, h1), O.conv(h1, o)) < 1e-12                                 # commutative
    assert O.conv(o, h1).sum() == pytest.approx(o.sum() * h1.sum(), rel=1e-11)            # DC gain = product of sums


@given(k=st.integers(1, 30), seed=st.integers(0, 2 ** 16))
@settings(**SETTINGS)
def test_zernike_sum_is_linear_and_orthonormal_basis(k, seed):
    rng = np.random.default_rng(seed)
    nms = [O.noll_to_nm(j) for j in range(1, k + 1)]
    x, y = O.make_xy_grid(24, diameter=2.0)
    c1, c2 = rng.standard_normal(k), rng.standard_normal(k)
    s = O.zernike_sum(c1 + 3 * c2, nms, x, y)
    assert rel_linf(s, O.zernike_sum(c1, nms, x, y) + 3 * O.zernike_sum(c2, nms, x, y)) < 1e-12
    r, t = O.cart_to_polar(x, y)
    Z = O.zernike_nm_seq(nms, r, t)
    assert rel_linf(np.tensordot(c1, Z, axes=(0, 0)), O.zernike_sum(c1, nms, x, y)) < 1e-12
