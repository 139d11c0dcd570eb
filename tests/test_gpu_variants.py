"""GPU: the opt-in kernel variants stay bit-identical to the defaults.

PB_FOCUS_V=3  -- the role-specialised single-kernel focus pipeline (csrc/focus_fused.cu) against the column + row
                 kernel pair on the same stack (same engine, same operation order: identical bits);
PB_MDFT_PAIR=1 -- the CTA-pair (cta_group::2) tcgen05 matrix DFT against the single-CTA kernel.
The switches are read once per process, so each variant runs in a subprocess (under a timeout: the fused kernel's
workers spin on each other's counters)."""
import hashlib
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import hashlib, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
import prysm_b200 as pb
from prysm_b200 import _ops, fttools as F
pb.config.precision = 32
gen = torch.Generator(device='cuda').manual_seed(11)
def rnd(*s): return torch.complex(torch.randn(s, generator=gen, device='cuda'), torch.randn(s, generator=gen, device='cuda'))
h = hashlib.sha1()
for n, b in ((1024, 5), (2048, 3)):
    stack = rnd(b, n, n)
    for d in (-1, 1):
        out = _ops.fft2_batch(stack, (2 * n, 2 * n), dir=d, scale=1.0 / (2 * n), shift_in=True, shift_out=True)
        h.update(out.cpu().numpy().tobytes())
print('focus', h.hexdigest())
n, m = 512, 256
x = (np.arange(n) - n // 2) * 0.1
f = (np.arange(m) - m // 2) * (0.37 / (n * 0.1))
ex = F.MDFT(x, x, f, f, -1, 0.5)
assert ex._tc is not None
print('mdft', hashlib.sha1(ex(rnd(n, n)).cpu().numpy().tobytes()).hexdigest())
'''


def _run(env):
    e = dict(os.environ)
    for k in ('PB_FOCUS_V', 'PB_MDFT_PAIR'):
        e.pop(k, None)
    e.update(env)
    r = subprocess.run([sys.executable, '-c', _SCRIPT % {'root': ROOT}], env=e, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    return dict(line.split() for line in r.stdout.strip().splitlines() if line.split()[0] in ('focus', 'mdft'))


@pytest.mark.gpu
def test_opt_in_variants_are_bit_identical():
    base = _run({})
    variant = _run({'PB_FOCUS_V': '3', 'PB_MDFT_PAIR': '1'})
    assert variant['focus'] == base['focus']
    assert variant['mdft'] == base['mdft']
