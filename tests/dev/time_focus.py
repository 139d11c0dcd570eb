"""GPU: time the batched focus call only (no oracle import): used under `ncu --metrics gpu__time_duration.sum`
for per-kernel launch lists of the variants selected by environment switches, and bare for CUDA-event timing."""
import os, sys, subprocess, threading, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import prysm_b200 as pb
from prysm_b200 import _ops

N, K, B = 2048, 4096, 16
reps = int(os.environ.get('REPS', '10'))
gen = torch.Generator(device='cuda').manual_seed(1)
stack = torch.complex(torch.randn((B, N, N), generator=gen, device='cuda'), torch.randn((B, N, N), generator=gen, device='cuda'))
out = torch.empty((B, K, K), dtype=torch.complex64, device='cuda')
def step():
    _ops.fft2_batch(stack, (K, K), dir=-1, scale=1.0 / K, shift_in=True, shift_out=True, out=out)
for _ in range(3): step()
torch.cuda.synchronize()
clk = []
def sample():
    try:
        r = subprocess.run(['nvidia-smi', '--query-gpu=clocks.sm,power.draw', '--format=csv,noheader,nounits', '-i', '0'],
                           capture_output=True, text=True, timeout=5)
        clk.append(r.stdout.strip())
    except Exception as e:
        clk.append(repr(e))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
th = threading.Thread(target=sample); 
e0.record()
for i in range(reps):
    step()
    if i == reps // 2: th.start()
e1.record(); torch.cuda.synchronize(); th.join()
us = e0.elapsed_time(e1) * 1e3 / (reps * B)
env = {k: v for k, v in os.environ.items() if k.startswith('PB_')}
print(f'{env}: {us:.1f} us/propagation frac {167772160 / us / 1e3 / 6571.2:.3f}  clocks(sm MHz, W)={clk}', flush=True)
