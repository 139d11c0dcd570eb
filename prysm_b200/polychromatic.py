"""Incoherent (polychromatic / multi-field) sums sharded over GPUs.

The reference has no multi-device code: its documented recipe is a Python loop over wavelengths
followed by `sum_of_2d_modes(stack, weights)` (docs/source/how-tos/Polychromatic Propagation.ipynb:86-98,
prysm/polynomials/fitting.py:37), and its docs advise users to parallelise that loop themselves.
Here the loop is the sharded unit of work:

    rank r of W takes units r, r+W, r+2W, ...         (no data-path communication)
    each unit's |field|^2 is accumulated, weighted, into ONE local fp32/fp64 plane as it is produced
    one sum-reduce of that plane (NCCL over NVLink when W > 1) gives the total on `dst` (or on all ranks)

A single 2-D field is never split across devices.
"""
import numpy as np
import torch
import torch.distributed as dist

from . import _ops
from . import propagation as P
from .conf import config


_SIDE = {}


def _side_stream(device):
    """One helper stream per device for plan construction (kept for the life of the process: its engine handle owns the
    twiddle tables of the single-line transforms the plans need)."""
    key = device.index if device.index is not None else torch.cuda.current_device()
    if key not in _SIDE:
        _SIDE[key] = torch.cuda.Stream(device=device)
    return _SIDE[key]


def _keep_alive(executor, stream):
    """The executor's plan tensors were allocated on the side stream and are read by kernels queued on `stream`: tell the
    caching allocator, so that freeing the executor cannot hand their memory out while those kernels are pending."""
    for v in vars(executor).values():
        for t in (v if isinstance(v, (list, tuple)) else (v,)):
            for u in (t if isinstance(t, (list, tuple)) else (t,)):
                if isinstance(u, torch.Tensor) and u.is_cuda:
                    u.record_stream(stream)


def shard_units(n_units, rank, world):
    """Indices of the units owned by `rank`: round-robin, so equal-cost units balance to within one."""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    return list(range(rank, n_units, world))


def _world(group):
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def sharded_incoherent_sum(n_units, accumulate_unit, plane, group=None, dst=None, shard=True):
    """Generic driver.  `accumulate_unit(i, plane)` must add unit i's weighted intensity into `plane`
    (a tensor on this rank's device).  After the local loop the planes are summed across ranks:
    onto `dst` if given (other ranks' planes are left as their partial sums), else onto every rank.
    shard=False runs every unit on the calling rank with no exchange (the single-GPU form inside a multi-rank job)."""
    rank, world = _world(group) if shard else (0, 1)
    for i in shard_units(n_units, rank, world):
        accumulate_unit(i, plane)
    if world > 1:
        if dst is None:
            dist.all_reduce(plane, op=dist.ReduceOp.SUM, group=group)
        else:
            dist.reduce(plane, dst=dst, op=dist.ReduceOp.SUM, group=group)
    return plane


def _native_czt_units(kind, opd, mine, wavelengths, weights, dx, efl, focal_dx, focal_samples, shift):
    """(K, units) for pb_polychromatic_czt -- units[i] = (kscale, shift, alpha, xc, f0, df, norm, weight) of this rank's i-th
    wavelength -- or None when the configuration is not the one that entry point covers: CZT, square pupil and focal
    grid, one Bluestein plan serving both axes, data in the configured precision."""
    import os
    from .fttools import czt_axis_scalars
    if kind != 'czt' or os.environ.get('PB_POLY_NATIVE', '1') == '0':
        return None
    if opd.ndim != 2 or opd.shape[0] != opd.shape[1] or focal_samples[0] != focal_samples[1] or opd.dtype != config.real_dtype:
        return None
    rows, K = [], None
    for i in mine:
        w = float(wavelengths[i])
        x, y, fx, fy = P.coordinates_for_focus(dx, tuple(opd.shape), focal_dx, focal_samples, w, efl, shift, dtype=np.float64)
        sx, sy = czt_axis_scalars(x, fx), czt_axis_scalars(y, fy)
        if len(sx) != 1 or sx != sy:
            return None
        _, (n, m, k, shf, alpha, _, xc, f0, df) = sx[0]
        if K is None:
            K = k
        if k != K:
            return None
        rows.append((P.phase_prefix(w).imag, shf, alpha, xc, f0, df, (dx * focal_dx) / (w * efl), float(weights[i])))
    return (K if K is not None else 0), np.asarray(rows, dtype=np.float64).reshape(-1, 8)


def polychromatic_psf(amplitude, phase, wavelengths, weights, dx, efl, focal_dx, focal_samples, kind='czt',
                      shift=(0, 0), group=None, dst=None, shard=True):
    """Weighted incoherent PSF over `wavelengths` on a common focal grid.

    Per wavelength (exactly the reference recipe): from_amp_and_phase -> prepare_executor(kind) ->
    focus_dft -> intensity; the weighted sum is accumulated as the planes are produced.  Returns the
    (focal_samples, focal_samples) real plane; with torch.distributed initialised the wavelengths are
    sharded over the ranks and the result is reduced (see sharded_incoherent_sum)."""
    amp = None if amplitude is None else _ops.asdevice(amplitude)
    opd = _ops.asdevice(phase)
    if opd.dtype not in (torch.float32, torch.float64):
        opd = opd.to(config.real_dtype)
    wavelengths = np.asarray(wavelengths, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    if wavelengths.shape != weights.shape:
        raise ValueError('one weight per wavelength is required')
    if isinstance(focal_samples, int):
        focal_samples = (focal_samples, focal_samples)
    # The executors compute in config.precision whatever the OPD's storage type is (`_prep` casts the field to
    # config.complex_dtype), so the accumulation plane is allocated in THAT real type -- never in opd.dtype -- and
    # `_ops.intensity` refuses a plane whose type does not match the field it is handed.
    plane = torch.zeros(tuple(focal_samples), dtype=config.real_dtype, device=opd.device)
    cplx_of_plane = torch.complex64 if plane.dtype == torch.float32 else torch.complex128

    rank, world = _world(group) if shard else (0, 1)
    mine = shard_units(len(wavelengths), rank, world)

    # CZT on a square, centred geometry (every BASELINE configuration): the whole loop of this rank is ONE native call
    # (pb_polychromatic_czt) -- per wavelength six scalars instead of ~7 library calls, two temporaries and an event, so the
    # loop is bound by its ~220 us of kernels per wavelength and not by the host's launch rate (203 ... 328 us per wavelength
    # were measured through the Python loop below on two boxes).  PB_POLY_NATIVE=0 selects the Python loop.
    units = _native_czt_units(kind, opd, mine, wavelengths, weights, dx, efl, focal_dx, focal_samples, shift)
    if units is not None:
        if mine:
            _ops.polychromatic_czt(amp, opd, focal_samples[0], units[0], units[1], plane)
        return sharded_incoherent_sum(0, None, plane, group=group, dst=dst, shard=shard)

    # Each wavelength has its own executor (the chirp rate dx*dfx/(wvl*efl) changes), built on the device from scalars in
    # three small, latency-bound launches (~20 us).  They are issued on a SIDE stream one unit ahead, so the next
    # wavelength's plan is ready when the current wavelength's transforms finish instead of sitting between them.
    main = torch.cuda.current_stream(opd.device)
    side = _side_stream(opd.device)
    shape = tuple(opd.shape)

    def build(i):
        with torch.cuda.stream(side):     # depends on nothing queued on the main stream
            ex = P.prepare_executor(dx, shape, focal_dx, focal_samples, float(wavelengths[i]), efl, shift, kind)
            ev = torch.cuda.Event()
            ev.record(side)
        return ex, ev
    pending = {mine[0]: build(mine[0])} if mine else {}

    def unit(i, acc):
        ex, ev = pending.pop(i)
        main.wait_event(ev)
        nxt = mine.index(i) + 1
        if nxt < len(mine):
            pending[mine[nxt]] = build(mine[nxt])
        wf = P.Wavefront.from_amp_and_phase(amp, opd, float(wavelengths[i]), dx)
        if hasattr(ex, 'intensity'):      # CZT: weight * |.|^2 is added to the plane by the last pass itself
            ex.intensity(wf.data, weight=float(weights[i]), out=acc)
        else:
            field = wf.focus_dft(ex).data
            if field.dtype != cplx_of_plane:
                field = field.to(cplx_of_plane)
            _ops.intensity(field, weight=float(weights[i]), out=acc)
        _keep_alive(ex, main)

    return sharded_incoherent_sum(len(wavelengths), unit, plane, group=group, dst=dst, shard=shard)


def polychromatic_psf_fft(amplitude, phase, wavelengths, weights, Q=2, group=None, dst=None):
    """Same sum on each wavelength's own FFT grid (`Wavefront.focus(Q)`): the synthesis, the padded FFT
    and the weighted |.|^2 accumulate are one fused call per wavelength (no complex field is written)."""
    amp = None if amplitude is None else _ops.asdevice(amplitude)
    opd = _ops.asdevice(phase)
    if opd.dtype not in (torch.float32, torch.float64):
        opd = opd.to(config.real_dtype)
    wavelengths = np.asarray(wavelengths, dtype=np.float64)
    weights = np.asarray(weights, dtype=np.float64)
    ky, kx = P._padded_shape(opd.shape, Q)
    plane = torch.zeros((ky, kx), dtype=opd.dtype, device=opd.device)

    def unit(i, acc):
        P.psf_from_amp_and_phase(amp, opd, float(wavelengths[i]), Q, weight=float(weights[i]), out=acc)

    return sharded_incoherent_sum(len(wavelengths), unit, plane, group=group, dst=dst)
