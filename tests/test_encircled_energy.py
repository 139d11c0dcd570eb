"""Baliga & Cohn encircled energy (reference prysm/otf.py:319-414): oracle vs the reference's stored values
(CPU) and the CUDA reduction vs both (GPU)."""
import numpy as np
import pytest

import prysm_oracle as O


def test_oracle_matches_reference_golden(small):
    I, pdx = small['wf_psf_intensity'], float(small['wf_psf_dx'])
    got = O.encircled_energy(I, pdx, small['wf_ee_radii'])
    assert np.allclose(got, small['wf_ee'], rtol=1e-12)
    assert O.encircled_energy(I, pdx, 5.0) == pytest.approx(float(small['wf_ee'][1]), rel=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('prec', [64, 32])
def test_gpu_encircled_energy(small, prec):
    import torch
    if not torch.cuda.is_available():
        pytest.skip('needs a CUDA device')
    import prysm_b200 as pb
    rdt = np.float64 if prec == 64 else np.float32
    I, pdx = small['wf_psf_intensity'].astype(rdt), float(small['wf_psf_dx'])
    radii = small['wf_ee_radii']
    got = pb.otf.encircled_energy(I, pdx, radii)
    tol = 1e-10 if prec == 64 else 5e-6
    assert np.allclose(got, small['wf_ee'], rtol=tol)
    one, data = pb.otf.encircled_energy(I, pdx, 5.0, return_more=True)
    assert isinstance(one, float) and one == pytest.approx(float(small['wf_ee'][1]), rel=tol)
    assert tuple(data.shape) == I.shape and data.is_complex()
    # monotone in radius and bounded by the total energy fraction
    assert np.all(np.diff(got) > 0) and got[-1] < 1.01
