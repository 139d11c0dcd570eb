"""CPU: the oracle's pupil-synthesis restatement (grids, masks, Jacobi / Zernike by recurrence)
against the reference's own outputs in tests/golden/synthesis.npz (oracle/make_golden.py synthesis)."""
import numpy as np
import pytest

import prysm_oracle as O
from conftest import rel_linf, load_golden


@pytest.fixture(scope='module')
def gold():
    return load_golden('synthesis.npz')


def test_grids_and_masks(gold):
    g = gold
    x, y = O.make_xy_grid((33, 40), dx=0.06)
    assert (x == g['grid_x']).all() and (y == g['grid_y']).all()
    r, t = O.cart_to_polar(x, y)
    assert rel_linf(r, g['grid_r']) < 1e-15 and rel_linf(t, g['grid_t']) < 1e-15
    assert (O.circle(1.0, r) == g['circle']).all()
    assert rel_linf(O.antialias(r - 1.0, 0.06), g['grey']) < 1e-15
    xd, yd = O.make_xy_grid(32, diameter=2.0)
    assert (xd == g['grid_xd']).all() and (yd == g['grid_yd']).all()


@pytest.mark.parametrize('i', range(4))
def test_jacobi_seq(gold, i):
    al, be = gold[f'jac{i}_ab']
    got = O.jacobi_seq([0, 1, 2, 5, 9, 14], al, be, gold['jac_x'])
    assert rel_linf(got, gold[f'jac{i}']) < 1e-14
    assert rel_linf(O.jacobi(9, al, be, gold['jac_x']), gold[f'jac{i}'][4]) < 1e-14


def test_zernike_seq_and_sum(gold):
    g = gold
    nms = [tuple(int(v) for v in nm) for nm in g['z_nms']]
    assert nms == [O.noll_to_nm(j) for j in range(1, 38)]
    r, t = g['grid_r'] / 1.2, g['grid_t']
    assert rel_linf(O.zernike_nm_seq(nms, r, t, True), g['z_seq_norm']) < 1e-13
    assert rel_linf(O.zernike_nm_seq(nms, r, t, False), g['z_seq_raw']) < 1e-13
    assert rel_linf(O.zernike_nm_seq([(5, -3)], r, t)[0], g['z_single']) < 1e-13
    # the closed-form radial polynomial used by the seeded BASELINE pupil agrees with the recurrence
    assert rel_linf(O.zernike_nm(5, -3, r, t), g['z_single']) < 1e-12
    assert rel_linf(O.zernike_sum(g['z_coefs'], nms, g['grid_x'] / 1.2, g['grid_y'] / 1.2), g['z_sum']) < 1e-13
    assert not O.zernike_sum([], [], g['grid_x'], g['grid_y']).any()
