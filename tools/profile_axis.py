"""One C5 plane (screened free_space at 4096^2), one C4 wavelength (CZT 2048^2 -> 2048^2, K = 4096) and one final CZT
focus 4096^2 -> 512^2: the launch sequence to capture with ncu (-k regex:axis_)."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prysm_b200 as pb
from prysm_b200 import propagation as P
from prysm_b200.polychromatic import polychromatic_psf

pb.config.precision = 32
gen = torch.Generator(device='cuda').manual_seed(1)
def crand(n): return torch.complex(torch.randn((n, n), generator=gen, device='cuda'), torch.randn((n, n), generator=gen, device='cuda'))
a4, scr = crand(4096), crand(4096)
wf = P.Wavefront(a4, 0.6328, 10.0 / 4096)
ex = wf.prepare_executor(100.0, 0.6328 * 10.0 / 4, 512, kind='czt')
N = 2048
amp = torch.ones((N, N), device='cuda'); opd = torch.randn((N, N), generator=gen, device='cuda') * 50
for rep in range(int(os.environ.get('REPS', '2'))):
    out = (wf * scr).free_space(dz=5.0, Q=1)
    psf = out.focus_dft(ex)
    poly = polychromatic_psf(amp, opd, [0.55], [1.0], 10.0 / N, 100.0, 2.5, N, kind='czt')
torch.cuda.synchronize()
print('done', float(psf.data.abs().max()), float(poly.max()))
