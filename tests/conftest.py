import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'oracle')):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def small():
    return np.load(os.path.join(GOLDEN, 'small.npz'))


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name))


def rel_linf(a, b):
    """max|a-b| / max|b| -- the parity metric of SURVEY.md section 8(d)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = float(np.abs(b).max())
    return float(np.abs(a - b).max()) / (den if den > 0 else 1.0)
