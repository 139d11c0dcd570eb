"""GPU diagnostic: accuracy of the library's c64 FFT passes vs an fp64 FFT (and cuFFT's c64 for scale)."""
import sys, os
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import prysm_b200 as pb
from prysm_b200 import _ops

def rel(a, b):
    return float(np.abs(a - b).max() / np.abs(b).max())

rng = np.random.default_rng(0)
for n in (16, 64, 128, 256, 1024, 4096):
    a = (rng.standard_normal((32, n)) + 1j * rng.standard_normal((32, n))).astype(np.complex64)
    ref = np.fft.fft(a.astype(np.complex128), axis=1)
    d = pb.asdevice(a)
    mine = _ops.fft1(d, n, axis=1).cpu().numpy()
    cu = torch.fft.fft(d, dim=1).cpu().numpy()
    at = np.ascontiguousarray(a.T)
    minec = _ops.fft1(pb.asdevice(at), n, axis=0).cpu().numpy()
    inv = _ops.fft1(d, n, axis=1, dir=+1, scale=1.0 / n).cpu().numpy()
    print(f'n={n:5d} rows {rel(mine, ref):.2e}  cols {rel(minec, ref.T):.2e}  inverse {rel(inv, np.fft.ifft(a.astype(np.complex128), axis=1)):.2e}  cuFFT {rel(cu, ref):.2e}')
