"""CPU: host-side logic of the mirror API that needs no kernel launch -- argument checks,
error types / messages, shape and sample-spacing bookkeeping, executor planning."""
import math
import os

import numpy as np
import pytest

from prysm_b200 import propagation, fttools
from prysm_b200.conf import config, Config


def test_config_precision_like_reference():
    c = Config()
    assert c.precision is np.float64 and c.precision_complex is np.complex128
    c.precision = 32
    assert c.precision is np.float32 and c.precision_complex is np.complex64
    c.precision = 'float64'
    assert c.precision is np.float64
    with pytest.raises(ValueError):
        c.precision = 'int32'
    with pytest.raises(ValueError):
        c.precision = 'not a dtype'


def test_sample_spacing_helpers():
    assert propagation.pupil_sample_to_psf_sample(10 / 2048, 4096, 0.6328, 100.0) == pytest.approx(3.164)
    assert propagation.psf_sample_to_pupil_sample(3.164, 4096, 0.6328, 100.0) == pytest.approx(10 / 2048)
    assert propagation.Q_for_sampling(10.0, 100.0, 0.5, 2.5) == pytest.approx(2.0)
    assert propagation.phase_prefix(0.5) == pytest.approx(1j * 2 * np.pi / 0.5 / 1e3)
    fdx, n = propagation.unit_cell_focal_grid(0.1, 10.0, 0.6328, 100.0, Q=2)
    assert n == 200 and fdx == pytest.approx(0.6328 * 100 / 0.1 / 200)


def test_padded_shapes_follow_pad2d_rule():
    assert propagation._padded_shape((9, 12), 1.5) == (14, 18)
    assert propagation._padded_shape((7, 9), 2) == (14, 18)
    assert propagation._padded_shape((8, 8), 1) == (8, 8)
    assert propagation._shape_before_pad((14, 18), 1.5) == (9, 12)


def test_coordinates_for_focus_matches_oracle():
    import prysm_oracle as O
    config.precision = 64
    got = propagation.coordinates_for_focus(0.1, (9, 12), 1.7, (8, 11), 0.6328, 100.0, (3.0, -2.0))
    want = O.coordinates_for_focus(0.1, (9, 12), 1.7, (8, 11), 0.6328, 100.0, (3.0, -2.0))
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


def test_next_fast_len_is_power_of_two():
    for n, want in ((1, 1), (2, 2), (3, 4), (4095, 4096), (4607, 8192), (4096, 4096)):
        assert fttools.next_fast_len(n) == want


def test_fftdft_validation_messages():
    config.precision = 64
    with pytest.raises(ValueError, match='uniformly spaced'):
        fttools._uniform_spacing(np.array([0.0, 1.0, 2.5]), 'x')
    with pytest.raises(ValueError, match='at least two samples'):
        fttools._uniform_spacing(np.array([0.0]), 'x')
    with pytest.raises(ValueError, match='not FFT-compatible'):
        fttools._fft_compatible_length(1 / 10.3, 8, 8, 'x/fx')
    with pytest.raises(ValueError, match='smaller than input/output'):
        fttools._fft_compatible_length(1 / 8, 16, 8, 'x/fx')
    assert fttools._fft_compatible_length(-1 / 32, 16, 32, 'x/fx') == 32


def test_prepare_executor_bad_kind():
    with pytest.raises(ValueError, match="kind must be 'mdft', 'czt', or 'fftdft'"):
        propagation.prepare_executor(0.1, 8, 1.0, 8, 0.5, 100.0, kind='nope')


def test_czt_axis_plan_matches_oracle_basis():
    """The fp64 host plan (chirps, kernel spectrum) equals the oracle's _czt_axis_basis up to the
    different internal K (any K >= N+M-1 gives the same transform)."""
    import prysm_oracle as O
    N, M, shift, alpha = 12, 11, 1.25, 0.013
    ax = fttools._CztAxis(N, M, shift, alpha, -1, 0.0, np.zeros(M))
    H, b, a = O._czt_axis_basis(N, M, ax.K, shift, alpha, np.float64, np.complex128, -1)
    assert np.allclose(ax.b, b, atol=1e-14) and np.allclose(ax.post, a, atol=1e-14)
    assert np.allclose(ax.H, H, atol=1e-12)
    assert ax.K == 32 and math.log2(ax.K).is_integer()


def test_widened_rows_host_checks_need_no_gpu():
    """Argument checks of the SURVEY 8(f) rows that fire before any kernel launch (reference messages)."""
    from prysm_b200 import coronagraph, polynomials
    with pytest.raises(TypeError, match='charge must be an integer'):
        coronagraph.vortex_phase_mask(2.5)
    assert callable(coronagraph.vortex_phase_mask(np.int64(2)))
    with pytest.raises(ValueError, match='requires field_at_fpm'):
        coronagraph.to_fpm_and_back_adjoint(None, None, None, return_fpm_grad=True)
    with pytest.raises(ValueError, match='requires field_at_fpm'):
        coronagraph.to_fpm_and_back_multiresolution_adjoint(None, None, None, return_fpm_grad=True)
    with pytest.raises(ValueError, match='requires field_at_lyot'):
        coronagraph.babinet_adjoint(None, None, None, None, return_lyot_grad=True)
    with pytest.raises(ValueError, match='two positive values'):
        fttools.fourier_resample(np.ones((4, 4)), (1, -1))
    a = np.ones((4, 4))
    assert fttools.fourier_resample(a, 1) is a
    import prysm_oracle as O
    assert [polynomials.noll_to_nm(j) for j in range(1, 80)] == [O.noll_to_nm(j) for j in range(1, 80)]
    assert polynomials.fringe_to_nm(1) == (0, 0) and polynomials.fringe_to_nm(4) == (2, 0)
    assert polynomials.zernike_norm(4, 0) == pytest.approx(math.sqrt(5))
    assert propagation.MultiResolutionExecutor([1, 2], [3, 4], [5, 6], [7, 8]).__len__() == 2


# ------------------------------------------------------------------------------------------
# round-2 host logic: validation of caller-supplied outputs, the bounded handle cache
# ------------------------------------------------------------------------------------------
def test_check_out_rejects_what_the_kernel_would_misread():
    """A user `out=` reaches the kernel as a raw pointer + pitch: wrong dtype / shape / device / stride must raise
    instead of writing out of bounds.  Pure host logic: CPU tensors suffice."""
    import torch
    from prysm_b200 import _ops
    ok = torch.empty((4, 6), dtype=torch.float32)
    assert _ops._check_out(ok, (4, 6), torch.float32, ok.device) is ok
    with pytest.raises(ValueError, match='dtype'):
        _ops._check_out(torch.empty((4, 6), dtype=torch.float64), (4, 6), torch.float32, ok.device)
    with pytest.raises(ValueError, match='shape'):
        _ops._check_out(torch.empty((4, 5), dtype=torch.float32), (4, 6), torch.float32, ok.device)
    with pytest.raises(ValueError, match='stride'):
        _ops._check_out(torch.empty((6, 4), dtype=torch.float32).t(), (4, 6), torch.float32, ok.device)
    with pytest.raises(ValueError, match='is on'):
        _ops._check_out(ok, (4, 6), torch.float32, torch.device('meta'))
    with pytest.raises(TypeError):
        _ops._check_out(np.zeros((4, 6), dtype=np.float32), (4, 6), torch.float32, ok.device)


def test_handle_cache_is_bounded_and_pins_graph_streams(monkeypatch):
    """One engine handle per (device, stream), at most MAX_HANDLES_PER_DEVICE un-pinned ones per device, least recently
    used first out; pinned handles (CUDA-graph capture streams) survive; release_handle closes on demand."""
    from prysm_b200 import _capi
    closed = []

    class Fake:
        def __init__(self, device):
            self.device = device

        def close(self):
            closed.append(self)
    monkeypatch.setattr(_capi, 'Handle', Fake)
    monkeypatch.setattr(_capi, '_handles', {})
    monkeypatch.setattr(_capi, '_pinned', set())
    cap = _capi.MAX_HANDLES_PER_DEVICE
    first = _capi.handle_for(0, 100)
    _capi.pin_handle(0, 100)
    hs = [_capi.handle_for(0, 200 + i) for i in range(cap)]
    assert _capi.handle_for(0, 200) is hs[0]                      # a hit refreshes its recency
    _capi.handle_for(0, 999)                                      # one too many un-pinned handles on device 0
    assert closed == [hs[1]] and (0, 201) not in _capi._handles   # the least recently used un-pinned one went
    assert _capi.handle_for(0, 100) is first and first not in closed
    _capi.handle_for(1, 5)                                        # another device has its own budget
    assert len(closed) == 1
    _capi.release_handle(0, 100)
    assert first in closed and (0, 100) not in _capi._handles and (0, 100) not in _capi._pinned


def test_native_polychromatic_units_host_logic():
    """polychromatic._native_czt_units: the records handed to pb_polychromatic_czt are the scalars the CZT executor itself
    would use (same helper), and every configuration the native loop does not cover falls back (None)."""
    import torch
    import prysm_b200 as pb
    from prysm_b200 import polychromatic as poly, propagation as P
    from prysm_b200.fttools import czt_axis_scalars
    pb.config.precision = 32
    try:
        N, M = 256, 128
        opd = torch.zeros((N, N), dtype=torch.float32)
        wv = np.array([0.5, 0.6, 0.7]); wt = np.array([0.2, 0.3, 0.5])
        K, units = poly._native_czt_units('czt', opd, [0, 2], wv, wt, 0.04, 100.0, 2.5, (M, M), (0, 0))
        assert K == 512 and units.shape == (2, 8) and units.dtype == np.float64
        for row, i in zip(units, (0, 2)):
            x, y, fx, fy = P.coordinates_for_focus(0.04, (N, N), 2.5, (M, M), float(wv[i]), 100.0, (0, 0), dtype=np.float64)
            (_, (n, m, k, shift, alpha, sign, xc, f0, df)), = czt_axis_scalars(x, fx)
            assert (n, m, k, sign) == (N, M, 512, -1)
            assert row[0] == pytest.approx(2 * np.pi / (wv[i] * 1e3))                 # OPD [nm] -> radians
            assert tuple(row[1:6]) == (shift, alpha, xc, f0, df)
            assert row[6] == pytest.approx(0.04 * 2.5 / (wv[i] * 100.0)) and row[7] == wt[i]
        assert poly._native_czt_units('czt', opd, [], wv, wt, 0.04, 100.0, 2.5, (M, M), (0, 0))[1].shape == (0, 8)
        # not covered: other executors, rectangular grids, an off-axis window (x and y plans differ), mixed precision,
        # a Bluestein length beyond the register engine (two half-length plans per axis), the opt-out switch
        assert poly._native_czt_units('mdft', opd, [0], wv, wt, 0.04, 100.0, 2.5, (M, M), (0, 0)) is None
        assert poly._native_czt_units('czt', opd, [0], wv, wt, 0.04, 100.0, 2.5, (M, M // 2), (0, 0)) is None
        assert poly._native_czt_units('czt', opd, [0], wv, wt, 0.04, 100.0, 2.5, (M, M), (10.0, 0.0)) is None
        assert poly._native_czt_units('czt', opd.double(), [0], wv, wt, 0.04, 100.0, 2.5, (M, M), (0, 0)) is None
        big = torch.zeros((4096, 4096), dtype=torch.float32)
        assert poly._native_czt_units('czt', big, [0], wv, wt, 0.0025, 100.0, 2.5, (512, 512), (0, 0)) is None
        os.environ['PB_POLY_NATIVE'] = '0'
        assert poly._native_czt_units('czt', opd, [0], wv, wt, 0.04, 100.0, 2.5, (M, M), (0, 0)) is None
    finally:
        os.environ.pop('PB_POLY_NATIVE', None)
        pb.config.precision = 64
