"""GPU: the fused role-specialised focus kernel (focus_fused.cu) against the two-kernel pipeline (bit-identical:
same engine, same operation order) on stacks of several sizes, plus its timing for the variant selected by the
PB_FUSED_* environment switches.  Run every variant in its own process under `timeout` (workers spin on counters)."""
import os, sys, subprocess, threading
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import prysm_b200 as pb
from prysm_b200 import _ops, _capi as capi

gen = torch.Generator(device='cuda').manual_seed(7)
def rnd(*shape):
    return torch.complex(torch.randn(shape, generator=gen, device='cuda'), torch.randn(shape, generator=gen, device='cuda'))

if '--check' in sys.argv:
    for N, Bs in ((2048, (2, 5, 16)), (1024, (3, 16))):
        K = 2 * N
        for B in Bs:
            stack = rnd(B, N, N)
            for d in (-1, 1):
                got = _ops.fft2_batch(stack, (K, K), dir=d, scale=1.0 / K, shift_in=True, shift_out=True)
                torch.cuda.synchronize()
                ok = True
                for i in range(B):
                    ref = _ops.fft2(stack[i], (K, K), dir=d, scale=1.0 / K, shift_in=True, shift_out=True)
                    if not torch.equal(got[i], ref):
                        ok = False
                        print(f'  MISMATCH N={N} B={B} dir={d} field {i}: max|diff| = {float((got[i] - ref).abs().max()):.3e} '
                              f'of {float(ref.abs().max()):.3e}', flush=True)
                print(f'N={N} B={B} dir={d}: fused == per-field two-kernel path: {ok}', flush=True)
            inten = _ops.fft2_batch(stack, (K, K), dir=-1, scale=1.0 / K, shift_in=True, shift_out=True, out_kind=capi.OUT_INTENSITY)
            ref = _ops.fft2(stack[B - 1], (K, K), dir=-1, scale=1.0 / K, shift_in=True, shift_out=True, out_kind=capi.OUT_INTENSITY)
            print(f'N={N} B={B} intensity: {bool(torch.equal(inten[B - 1], ref))}', flush=True)
    # repeated launches reuse the ring and the counters
    stack = rnd(16, 2048, 2048)
    a = _ops.fft2_batch(stack, (4096, 4096), dir=-1, scale=1.0 / 4096, shift_in=True, shift_out=True)
    for _ in range(5):
        b = _ops.fft2_batch(stack, (4096, 4096), dir=-1, scale=1.0 / 4096, shift_in=True, shift_out=True)
    print('repeatable:', bool(torch.equal(a, b)), flush=True)

N, K, B = 2048, 4096, int(os.environ.get('B', '16'))
reps = int(os.environ.get('REPS', '20'))
stack = rnd(B, N, N)
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
th = threading.Thread(target=sample)
e0.record()
for i in range(reps):
    step()
    if i == reps // 2: th.start()
e1.record(); torch.cuda.synchronize(); th.join()
us = e0.elapsed_time(e1) * 1e3 / (reps * B)
env = {k: v for k, v in os.environ.items() if k.startswith('PB_')}
print(f'{env} B={B}: {us:.1f} us/propagation frac {167772160 / us / 1e3 / 6571.2:.3f}  clocks(sm MHz, W)={clk}', flush=True)
