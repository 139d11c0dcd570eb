"""CPU: the C-ABI library builds for sm_100a, loads, and exports exactly the symbols
include/prysm_b200.h declares (no compute calls here -- there is no GPU on this box)."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, 'include', 'prysm_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(pb_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_header_symbol():
    from prysm_b200 import _capi
    lib = ctypes.CDLL(_capi.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), f'{n} declared in the header but not exported'
    assert sorted(_capi.SIGNATURES) == names, 'ctypes SIGNATURES and the header diverge'


def test_version_and_null_handle_are_safe():
    from prysm_b200 import _capi
    assert 'sm_100a' in _capi.version()
    assert _capi.lib.pb_last_error(None) == b'null handle'
    assert _capi.lib.pb_launch_count(None) == 0
    assert _capi.lib.pb_mdft_work_elems(512, 4096, 512, 4096, 0, 1) == 512 * 4096


def test_no_cpu_fallback_without_device():
    """On a box without a GPU the product must fail loudly, not silently compute on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    import numpy as np
    from prysm_b200 import propagation, _capi
    with pytest.raises(_capi.B200Error):
        propagation.focus(np.ones((8, 8), dtype=np.complex64), 2)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, 'prysm_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'prysm_oracle' not in txt and 'import oracle' not in txt, f
