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


def test_python_constants_match_the_header_enums():
    """The integer codes the ctypes layer passes are the enum values the header declares."""
    from prysm_b200 import _capi
    src = open(os.path.join(ROOT, 'include', 'prysm_b200.h')).read()
    enums = {k: int(v) for k, v in re.findall(r'\b(PB_[A-Z0-9_]+)\s*=\s*(-?\d+)', src)}
    pairs = {'PB_C64': _capi.PB_C64, 'PB_C128': _capi.PB_C128, 'PB_IN_COMPLEX': _capi.IN_COMPLEX, 'PB_IN_REAL': _capi.IN_REAL,
             'PB_IN_AMP_OPD': _capi.IN_AMP_OPD, 'PB_AMP_NONE': _capi.AMP_NONE, 'PB_AMP_REAL': _capi.AMP_REAL,
             'PB_AMP_U8': _capi.AMP_U8, 'PB_OUT_COMPLEX': _capi.OUT_COMPLEX, 'PB_OUT_INTENSITY': _capi.OUT_INTENSITY,
             'PB_OUT_ACCUMULATE': _capi.OUT_ACCUMULATE, 'PB_MASK_REAL': _capi.MASK_REAL, 'PB_MASK_COMPLEX': _capi.MASK_COMPLEX,
             'PB_MASK_CONJ': _capi.MASK_CONJ, 'PB_MASK_ONE_MINUS': _capi.MASK_ONE_MINUS, 'PB_MASK_REAL_OUT': _capi.MASK_REAL_OUT,
             'PB_MASK_ACCUMULATE': _capi.MASK_ACCUMULATE}
    for name, val in pairs.items():
        assert enums[name] == val, name
    assert enums['PB_OK'] == 0 and enums['PB_ERR_INVALID'] == -1      # Handle.check maps -1 to ValueError


def test_only_tests_smoke_and_bench_touch_the_oracle():
    """oracle/ is test infrastructure: nothing outside tests/, bench.py and __graft_entry__.py may import it."""
    allowed = {os.path.join(ROOT, 'bench.py'), os.path.join(ROOT, '__graft_entry__.py')}
    for dirpath, dirs, files in os.walk(ROOT):
        dirs[:] = [d for d in dirs if d not in ('.git', 'tests', 'oracle', 'gpurun_out', 'baseline', '__pycache__', '.pytest_cache')]
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith('.py') and path not in allowed:
                assert 'prysm_oracle' not in open(path).read(), path


def test_header_is_valid_c99_and_the_c_example_links():
    """The boundary is a C ABI: the header must compile as plain C (not only as C++), and the example that calls it
    from C must link against the built library."""
    import shutil
    import subprocess
    import tempfile
    from prysm_b200 import _capi
    gcc = shutil.which('gcc')
    if gcc is None:
        pytest.skip('no gcc on this box')
    src = os.path.join(ROOT, 'examples', 'focus_from_c.c')
    inc = os.path.join(ROOT, 'include')
    r = subprocess.run([gcc, '-std=c99', '-Wall', '-Wextra', '-pedantic', '-Werror', '-I', inc, '-fsyntax-only', src],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    cudart = '/usr/local/cuda/lib64'
    if os.path.isdir(cudart):
        with tempfile.TemporaryDirectory() as d:
            r = subprocess.run([gcc, '-std=c99', '-I', inc, src, '-L', os.path.dirname(_capi.LIB_PATH), '-lprysm_b200',
                                '-L', cudart, '-lcudart', '-o', os.path.join(d, 'focus_from_c')], capture_output=True, text=True)
            assert r.returncode == 0, r.stderr


def test_every_entry_point_rejects_a_null_handle():
    """No entry point may dereference a NULL handle: each returns PB_ERR_INVALID (-1) instead (nothing crosses the
    boundary as a crash).  Needs no GPU -- the check is the first thing every call does."""
    from prysm_b200 import _capi
    no_handle = {'pb_create', 'pb_destroy', 'pb_last_error', 'pb_version', 'pb_launch_count', 'pb_mdft_work_elems',
                 'pb_mdft_tc_supported', 'pb_mdft_tc_work_bytes', 'pb_polychromatic_czt_work_bytes'}
    checked = 0
    for name, (_, args) in _capi.SIGNATURES.items():
        if name in no_handle:
            continue
        vals = []
        for a in args:
            if a in (ctypes.c_int, ctypes.c_longlong):
                vals.append(0)
            elif a is ctypes.c_double:
                vals.append(0.0)
            else:
                vals.append(None)       # pointers, including the handle
        assert getattr(_capi.lib, name)(*vals) == -1, name
        checked += 1
    assert checked == len(_capi.SIGNATURES) - len(no_handle)
