"""GPU: the batched focus call (pb_fft2_batch) gives, field for field, exactly what the single-field call gives --
the same kernels index one more dimension -- for every group size, for ragged last groups, on the fused shapes and on
the generic fallback, and for the fused phase-screen input."""
import numpy as np
import pytest
import torch

import prysm_oracle as O
from conftest import rel_linf

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def pb():
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import prysm_b200
    return prysm_b200


def crand(shape, seed, dt=torch.float32):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.complex(torch.randn(shape, generator=g, device='cuda', dtype=dt), torch.randn(shape, generator=g, device='cuda', dtype=dt))


@pytest.mark.parametrize('n,batch', [(512, 1), (512, 5), (1024, 3), (2048, 9), (96, 4), (100, 3)])
def test_batched_equals_single(pb, n, batch):
    P = pb.propagation
    w = crand((batch, n, n), n + batch)
    got = P.focus(w, 2)                               # (B, 2n, 2n): fused kernels for 512/1024/2048, generic passes otherwise
    assert tuple(got.shape) == (batch, 2 * n, 2 * n)
    for b in range(batch):
        assert torch.equal(got[b], P.focus(w[b], 2)), b
    back = P.unfocus(w, 2)
    for b in (0, batch - 1):
        assert torch.equal(back[b], P.unfocus(w[b], 2))
    ref = O.focus(w[batch - 1].cpu().numpy().astype(np.complex128), 2)
    assert rel_linf(got[batch - 1].cpu().numpy(), ref) < 1e-6


def test_batched_complex128_and_strided_views(pb):
    P = pb.propagation
    w = crand((3, 64, 64), 7, torch.float64)
    got = P.focus(w, 2)
    for b in range(3):
        assert torch.equal(got[b], P.focus(w[b], 2))
    assert rel_linf(got[1].cpu().numpy(), O.focus(w[1].cpu().numpy(), 2)) < 1e-12


def test_batched_phase_screen_input_shared_amplitude(pb):
    """PB_IN_AMP_OPD through the batch entry point: one amplitude for all fields, one OPD per field."""
    from prysm_b200 import _ops
    n, batch = 512, 3
    g = torch.Generator(device='cuda').manual_seed(3)
    amp = torch.rand((n, n), generator=g, device='cuda') > 0.3
    opd = torch.randn((batch, n, n), generator=g, device='cuda') * 60.0
    k = 2 * np.pi / 0.6328 / 1e3
    got = _ops.fft2_batch(None, (2 * n, 2 * n), dir=-1, scale=1.0 / (2 * n), shift_in=True, shift_out=True, amp=amp, opd=opd, kscale=k)
    for b in range(batch):
        one = pb.propagation.psf_from_amp_and_phase(amp, opd[b], 0.6328, 2, field=True)
        assert torch.equal(got[b], one)
    amps = torch.stack([amp, ~amp, amp])
    got2 = _ops.fft2_batch(None, (2 * n, 2 * n), dir=-1, scale=1.0 / (2 * n), shift_in=True, shift_out=True, amp=amps, opd=opd, kscale=k)
    assert torch.equal(got2[1], pb.propagation.psf_from_amp_and_phase(~amp, opd[1], 0.6328, 2, field=True))


def test_wavefront_focus_on_a_stack(pb):
    """Wavefront.focus / unfocus accept a (B, N, N) stack: sample spacing from the last axis, fields batched."""
    P = pb.propagation
    w = crand((3, 512, 512), 11)
    wf = P.Wavefront(w, 0.6328, 10.0 / 512)
    psf = wf.focus(100.0, Q=2)
    one = P.Wavefront(w[1], 0.6328, 10.0 / 512).focus(100.0, Q=2)
    assert psf.dx == one.dx and psf.space == 'psf' and tuple(psf.data.shape) == (3, 1024, 1024)
    assert torch.equal(psf.data[1], one.data)
    back = psf.unfocus(100.0, Q=1)
    assert back.dx == pytest.approx(wf.dx) and tuple(back.data.shape) == (3, 1024, 1024)


def test_batched_fused_intensity(pb):
    P = pb.propagation
    w = crand((5, 512, 512), 21)
    got = P.focus_intensity(w, 2)
    assert got.dtype == torch.float32 and tuple(got.shape) == (5, 1024, 1024)
    for b in (0, 4):
        assert torch.equal(got[b], P.focus_intensity(w[b], 2))
    with pytest.raises(ValueError):
        P.focus_intensity(w, 2, weight=0.5)
