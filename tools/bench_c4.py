"""BASELINE configs[3] (C4): 2048x2048 x 64-wavelength polychromatic PSF, wavelengths sharded over the ranks,
one NCCL sum-reduce of the weighted intensity plane.  Run alone (1 GPU) or under torch.distributed.run.

    python tools/bench_c4.py
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/bench_c4.py

Per wavelength (the reference recipe, docs/source/how-tos/Polychromatic Propagation.ipynb:86-98):
from_amp_and_phase -> prepare_executor(kind='czt') (K = 4096) -> focus_dft -> weighted |.|^2 accumulate.
Timing: barrier + synchronize on both sides, CUDA events, max over ranks, the reduce inside the timed region.
"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

world = int(os.environ.get('WORLD_SIZE', '1'))
rank = int(os.environ.get('RANK', '0'))
local = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local)
dev = torch.device('cuda', local)
if world > 1:
    os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
    dist.init_process_group('nccl', device_id=dev)

import prysm_b200 as pb  # noqa: E402
from prysm_b200.polychromatic import polychromatic_psf  # noqa: E402

pb.config.precision = 32
N = M = 2048
dx = 10.0 / N
g = torch.Generator(device=dev).manual_seed(20260923)
yy, xx = torch.meshgrid(torch.arange(N, device=dev) - N // 2, torch.arange(N, device=dev) - N // 2, indexing='ij')
amp = (xx * xx + yy * yy) <= (N // 2) ** 2
opd = torch.randn((N, N), generator=g, device=dev) * 30.0
wvls = np.linspace(0.5, 0.7, 64)
wts = np.full(64, 1 / 64)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


def run():
    return polychromatic_psf(amp, opd, wvls, wts, dx, 100.0, 2.5, M, kind='czt', dst=0 if world > 1 else None)


for _ in range(3):
    run()
barrier()
reps = 5
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
barrier()
e0.record()
for _ in range(reps):
    out = run()
e1.record()
barrier()
ms = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
if rank == 0:
    t = float(ms.item())
    print(json.dumps({'workload': 'C4: 2048^2 x 64 wavelengths, CZT -> 2048^2, weighted incoherent sum, 1 NCCL reduce',
                      'n_gpus': world, 'ms_per_polychromatic_psf': t, 'psf_per_s': 1e3 / t,
                      'us_per_wavelength_per_gpu': t * 1e3 / (64 / world),
                      'algorithmic_GBps_per_gpu': 50331648 * (64 / world) / (t * 1e-3) / 1e9,
                      'checksum': float(out.double().sum())}), flush=True)
if world > 1:
    dist.destroy_process_group()
