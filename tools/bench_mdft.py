"""GPU timing + accuracy of the tensor-core matrix DFT at the C3 shape (4096^2 -> 512^2)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prysm_b200 as pb
from prysm_b200 import propagation as P, fttools as F

pb.config.precision = 32
n, m = 4096, 512
gen = torch.Generator(device='cuda').manual_seed(7)
a = torch.complex(torch.randn((n, n), generator=gen, device='cuda'), torch.randn((n, n), generator=gen, device='cuda'))
ex = P.prepare_executor(10.0 / n, (n, n), 0.6328 * 10.0 / 4, (m, m), 0.6328, 100.0, kind='mdft')
out = ex(a)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for _ in range(3): ex(a)
torch.cuda.synchronize()
e0.record()
for _ in range(20): ex(a)
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 1e3 / 20
flops = 8 * (m * n * n + m * n * m)
# accuracy against the CUDA-core complex128 GEMM path of the same executor geometry (fp64 arithmetic on the device)
pb.config.precision = 64
ex64 = P.prepare_executor(10.0 / n, (n, n), 0.6328 * 10.0 / 4, (m, m), 0.6328, 100.0, kind='mdft')
ref = ex64(a.to(torch.complex128))
err = float((out.to(torch.complex128) - ref).abs().max() / ref.abs().max())
env = {k: v for k, v in os.environ.items() if k.startswith('PB_')}
print(f'{env}: MDFT C3 {us:.1f} us/apply  {flops / us / 1e6:.1f} TFLOP/s algorithmic  frac of bf16 burst peak {flops / us / 1e6 / 1719.2:.3f}  rel L-inf vs fp64 {err:.2e}')
