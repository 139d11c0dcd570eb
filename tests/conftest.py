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


# ---- measured-vs-bound bookkeeping of the BASELINE-size parity tests -----------------------------------------------
# `within(name, value, bound)` asserts value < bound and remembers the pair; at the end of a session that recorded
# anything the table is written to gpurun_out/test_measurements.json (scratch), so that the bounds in the tests can be
# kept at the documented figure x <= 1.5 instead of drifting.
_MEASURED = {}


def within(name, value, bound):
    value = float(value)
    _MEASURED[name] = {'measured': value, 'bound': float(bound)}
    assert value < bound, f'{name}: {value:.3e} is not below {bound:.3e}'
    return value


def pytest_sessionfinish(session, exitstatus):
    if not _MEASURED:
        return
    import json
    out = os.path.join(ROOT, 'gpurun_out')
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, 'test_measurements.json'), 'w') as fh:
            json.dump(_MEASURED, fh, indent=1, sort_keys=True)
    except OSError:
        pass
