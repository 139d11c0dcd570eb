#!/usr/bin/env python
"""bench.py -- BASELINE.json metric: 2048x2048 pupil -> PSF propagations/sec.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE configs[1], SURVEY.md 8d "C2"): one unit = `Wavefront.focus(efl, Q=2)` of a
2048^2 complex64 Zernike-aberrated pupil -> 4096^2 complex64 field.  One step = one pass over a
batch of BATCH distinct pupils resident in HBM through ONE batched library call (BATCH*32 MiB of
inputs and BATCH*128 MiB of distinct outputs >> 126 MB L2), so no step can be served from cache.  Multi-GPU = independent replicas
with disjoint batches (weak scaling, no collective on the data path).

The JSON line carries: value (device-resident throughput, CUDA events, max over ranks), e2e (same
metric through the public API with pinned HOST buffers, H2D + D2H inside the timed region),
roofline (algorithmic bytes / measured duration vs the measured HBM peak), cpu_baseline (the
oracle port = the reference's numpy/scipy algorithm, timed on this box's host cores), clocks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N = 2048
Q = 2
K = N * Q
HENE = 0.6328
EFL = 100.0
BATCH = 16                       # pupils per step (16 x 32 MiB = 512 MiB of distinct inputs)
ALG_BYTES = 8 * N * N + 8 * K * K  # read pupil + write field = 167 772 160 B / propagation
METRIC = '2048x2048 pupil->PSF propagations/sec'


def oracle():
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import prysm_oracle as O
    return O


def make_pupils(count, seed0=20260923):
    """Seeded complex64 pupils: the SURVEY 8(d) aperture with per-pupil Zernike coefficients."""
    import numpy as np
    O = oracle()
    amp, opd, dx = O.synthetic_pupil(N, np.float32)
    base = O.from_amp_and_phase(amp, opd.astype(np.float64), HENE).astype(np.complex64)
    out = []
    rng = np.random.default_rng(seed0)
    for i in range(count):
        # distinct inputs: a per-pupil global piston + tilt keeps |P| in {0,1} and costs nothing to build
        ph = rng.uniform(0, 2 * np.pi)
        tilt = rng.uniform(-3, 3, 2)
        g = np.arange(N, dtype=np.float32) / N
        mod = np.exp(1j * (ph + 2 * np.pi * (tilt[0] * g[:, None] + tilt[1] * g[None, :]))).astype(np.complex64)
        out.append(base * mod)
    return out, dx


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--query-gpu={q}', '--format=csv,noheader,nounits',
                                          '-i', str(self.index), '-lms', '20'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(',')]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        rows = [r for (ts, r) in self.rows if t0 is None or t0 <= ts <= t1 + 0.05]
        window = 'timed region'
        if len(rows) < 2:   # nvidia-smi ticks are coarse against a ~100 ms region: use every sample under the same load
            rows, window = [r for (_, r) in self.rows], 'warm-up + timed region (same workload)'
        for r in rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith('active'):
                        reasons.add(nm)
            except Exception:
                pass
        sm.sort()
        med = sm[len(sm) // 2] if sm else None
        return {'sm_mhz': med, 'sm_max_mhz': mx, 'reasons': sorted(reasons), 'samples': len(sm), 'window': window}


def cpu_focus_rate(seconds_budget, workers, pupils):
    """Time the oracle port of Wavefront.focus(Q=2) on complex64 2048^2 pupils with scipy.fft workers."""
    import numpy as np
    from scipy import fft as sfft
    O = oracle()
    done, t_total = 0, 0.0
    with sfft.set_workers(workers):
        O.focus(pupils[0], Q)  # warm-up (plan caches, page faults)
        t_end = time.perf_counter() + seconds_budget
        while True:
            t0 = time.perf_counter()
            out = O.focus(pupils[done % len(pupils)], Q)
            t_total += time.perf_counter() - t0
            done += 1
            if time.perf_counter() > t_end and done >= 3:
                break
    assert out.dtype == np.complex64 and out.shape == (K, K)
    return done / t_total, done, t_total


def measure_mdft_c3(pb, peaks):
    """BASELINE configs[2]: 4096^2 -> 512^2 fixed-sampling focus via MDFT on the tcgen05 tensor cores.
    Algorithmic flops: 8*(My*Ny*Nx + My*Nx*Mx) = 77 309 411 328 per apply."""
    import torch
    from prysm_b200 import propagation as P
    n, m = 4096, 512
    gen = torch.Generator(device='cuda').manual_seed(7)
    a = torch.complex(torch.randn((n, n), generator=gen, device='cuda'), torch.randn((n, n), generator=gen, device='cuda'))
    ex = P.prepare_executor(10.0 / n, (n, n), HENE * 10.0 / 4, (m, m), HENE, EFL, kind='mdft')
    for _ in range(3):
        ex(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        ex(a)
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / reps
    flops = 8 * (m * n * n + m * n * m)
    peak = float(peaks.get('bf16_tflops', 1590.0))
    ach = flops / sec / 1e12
    return {'workload': 'C3: 4096x4096 complex64 -> 512x512 focus_dft(MDFT), 3xTF32 tcgen05 complex GEMM',
            'us_per_apply': sec * 1e6, 'applies_per_s': 1.0 / sec, 'tensor_core_path': ex._tc is not None,
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peak, 'unit': 'TFLOP/s', 'frac': ach / peak,
                         'traffic': None, 'peak_source': 'MEASURED_PEAKS.json bf16_tflops (measured, burst)',
                         'algorithmic_flops_per_apply': flops, 'issued_tf32_tflops': 3 * ach,
                         'note': 'algorithmic flops vs the bf16 peak; the path issues 3 TF32 MMAs per product (TF32 runs at half the bf16 rate)'}}


def measure_fused_psf(stack, peak):
    """SURVEY 8(d) 'fused PSF variant': unit = focus(...).intensity with |.|^2 formed in the last pass (fp32 out).
    Algorithmic bytes 8*N^2 + 4*K^2 = 100 663 296 per PSF."""
    import torch
    from prysm_b200 import _ops
    from prysm_b200._capi import OUT_INTENSITY
    nb = stack.shape[0]
    out = torch.empty((nb, K, K), dtype=torch.float32, device=stack.device)

    def step():
        _ops.fft2_batch(stack, (K, K), dir=-1, scale=1.0 / K, shift_in=True, shift_out=True, out_kind=OUT_INTENSITY, out=out)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20
    e0.record()
    for _ in range(reps):
        step()
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / (reps * nb)
    alg = 8 * N * N + 4 * K * K
    return {'workload': 'C2 fused PSF variant: 2048x2048 complex64 pupil -> |focus(Q=2)|^2 4096x4096 float32, batched',
            'us_per_psf': sec * 1e6, 'psf_per_s': 1.0 / sec,
            'roofline': {'bound': 'hbm', 'achieved': alg / sec / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': alg / sec / 1e9 / peak, 'traffic': None, 'algorithmic_bytes_per_psf': alg}}


def run_reference(args):
    """--impl reference: the reference's algorithm (oracle port: numpy + scipy.fft pocketfft, the same
    third-party FFT the reference calls) on this box's host cores, all threads, same config/metric."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    import numpy as np
    from scipy import fft as sfft
    O = oracle()
    cores = os.cpu_count() or 1
    per_step = 2                                   # bounded sample: 2 propagations per step
    pupils, _ = make_pupils(2)
    with sfft.set_workers(cores):
        for _ in range(max(1, args.warmup)):
            O.focus(pupils[0], Q)
        t0 = time.perf_counter()
        for s in range(args.steps):
            for i in range(per_step):
                out = O.focus(pupils[i % len(pupils)], Q)
        dt = time.perf_counter() - t0
    assert out.dtype == np.complex64
    value = args.steps * per_step / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': 'propagations/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': 1e3 * dt / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex64',
        'data': 'synthetic',
        'config': {'workload': 'C2: 2048x2048 complex64 pupil -> Wavefront.focus(Q=2) -> 4096x4096 field',
                   'propagations_per_step': per_step, 'host_threads': cores},
        'cpu_baseline': {'value': value, 'unit': 'propagations/s', 'cores': cores, 'kind': 'port',
                         'sample': f'{args.steps * per_step} propagations, scipy.fft workers={cores}'},
        'e2e': {'value': value, 'unit': 'propagations/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=dev)

    import prysm_b200 as pb
    from prysm_b200 import propagation as P, _ops
    pb.config.precision = 32

    # ---- synthetic inputs: BATCH distinct pupils per rank, resident in HBM before timing
    host_pupils, dx = make_pupils(2, seed0=20260923 + rank)
    base = [torch.from_numpy(p).to(dev) for p in host_pupils]
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    pupils = []
    for i in range(BATCH):  # distinct unit-modulus piston per pupil (built once, outside timing)
        ph = torch.rand((), generator=gen, device=dev) * 6.2831853
        pupils.append((base[i % 2] * torch.polar(torch.ones((), device=dev), ph)).contiguous())
    stack = torch.stack(pupils)                                   # (BATCH, N, N): 512 MiB of distinct inputs
    del pupils
    out = torch.empty((BATCH, K, K), dtype=torch.complex64, device=dev)   # 2 GiB of distinct outputs
    scale = 1.0 / K
    torch.cuda.synchronize()

    def step():  # one batched library call (pb_fft2_batch): the fields of a step share launches, 8 per launch pair
        _ops.fft2_batch(stack, (K, K), dir=-1, scale=scale, shift_in=True, shift_out=True, out=out)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(max(3, args.warmup)):
        step()
    barrier()
    l0 = _ops.launch_count(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start = time.perf_counter()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    t_end = time.perf_counter()
    ms = e0.elapsed_time(e1)
    launches = _ops.launch_count(dev) - l0
    clocks = sampler.stop(t_start, t_end) if rank == 0 else None
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    props = args.steps * BATCH
    value = world * props / (ms_max * 1e-3)

    # ---- e2e: public API, pinned host buffers in, pinned host buffers out, copies inside the timed region
    e2e_props = 8
    hin = [torch.from_numpy(host_pupils[i % 2]).pin_memory() for i in range(2)]
    hout = [torch.empty((K, K), dtype=torch.complex64).pin_memory() for _ in range(2)]
    streams = [torch.cuda.Stream(dev) for _ in range(2)]

    def e2e_pass(n):
        for i in range(n):
            s = streams[i % 2]
            with torch.cuda.stream(s):  # double-buffered: copy-in, propagate, copy-out per stream
                d = hin[i % 2].to(dev, non_blocking=True)
                wf = P.Wavefront(d, HENE, dx).focus(EFL, Q=Q)
                hout[i % 2].copy_(wf.data, non_blocking=True)
        for s in streams:
            s.synchronize()

    e2e_pass(2)
    barrier()
    t0 = time.perf_counter()
    e2e_pass(e2e_props)
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    t2 = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t2, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_props / float(t2.item())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        except Exception:
            pass
        peak = float(peaks.get('hbm_gbs', 6650.0))
        peak_src = 'MEASURED_PEAKS.json hbm_gbs (measured)' if 'hbm_gbs' in peaks else 'fallback 6650 GB/s'
        t_prop = ms_max * 1e-3 / props
        achieved = ALG_BYTES / t_prop / 1e9
        cpu = None
        if world == 1:
            cores = os.cpu_count() or 1
            rate, n_done, secs = cpu_focus_rate(12.0, cores, host_pupils)
            cpu = {'value': rate, 'unit': 'propagations/s', 'cores': cores, 'kind': 'port',
                   'sample': f'{n_done} propagations of the same workload in {secs:.1f} s, scipy.fft workers={cores}'}
        line = {
            'metric': METRIC, 'value': value, 'unit': 'propagations/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': max(3, args.warmup), 'ms_per_step': ms_max / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'complex64', 'data': 'synthetic',
            'config': {'workload': 'C2: 2048x2048 complex64 pupil -> Wavefront.focus(Q=2) -> 4096x4096 field',
                       'propagations_per_step': BATCH, 'parallelism': f'replicas x{world}',
                       'l2_policy': f'{BATCH} distinct 32 MiB inputs + {BATCH} distinct 128 MiB outputs per step (>> 126 MB L2)',
                       'call': 'one pb_fft2_batch per step (propagation.focus on a (16, 2048, 2048) stack)'},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s', 'frac': achieved / peak,
                         # (dram__bytes_read + dram__bytes_write) / 8 fields of the two kernels of one launch pair from the
                         # committed `ncu --set full` capture of this command (profiles/r01_ncu_focus_batched_summary.txt):
                         # 33.6 + 60.8 (column kernel) + 67.1 + 126.9 (row kernel) MB -- with 8 fields per launch pair
                         # the 64 MiB intermediates spill to HBM once each way
                         'traffic': (268553216 + 486104000 + 536974336 + 1014972000) // 8, 'peak_source': peak_src,
                         'algorithmic_bytes_per_propagation': ALG_BYTES,
                         'kernel': 'fused focus pipeline (all passes of one propagation), per GPU',
                         'us_per_propagation': t_prop * 1e6},
            'e2e': {'value': e2e_value, 'unit': 'propagations/s', 'h2d_bytes_per_step': 8 * N * N,
                    'd2h_bytes_per_step': 8 * K * K,
                    'note': 'per propagation: pinned host pupil in, 4096^2 complex64 field back to pinned host memory'},
            'gpu_launches': launches,
            'clocks': clocks,
        }
        if cpu is not None:
            line['cpu_baseline'] = cpu
        if world == 1:
            line['mdft_c3'] = measure_mdft_c3(pb, peaks)
            try:
                line['fused_psf'] = measure_fused_psf(stack, peak)
            except Exception as exc:  # an extra: never take the headline line down with it
                line['fused_psf'] = {'error': repr(exc)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()
