"""ctypes binding of libprysm_b200.so (the C ABI declared in include/prysm_b200.h).

The library is dlopen'ed directly -- no pybind / torch-extension layer.  If it is missing the
import fails loudly: there is no CPU or PyTorch fallback anywhere in this package.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, '_lib', 'libprysm_b200.so')

PB_C64, PB_C128 = 0, 1
IN_COMPLEX, IN_REAL, IN_AMP_OPD = 0, 1, 2
AMP_NONE, AMP_REAL, AMP_U8 = 0, 1, 2
OUT_COMPLEX, OUT_INTENSITY, OUT_ACCUMULATE = 0, 1, 2
OP_N, OP_T, OP_H, OP_C = 0, 1, 2, 3
MASK_REAL, MASK_COMPLEX = 0, 1
MASK_CONJ, MASK_ONE_MINUS, MASK_REAL_OUT, MASK_ACCUMULATE = 1, 2, 4, 8

_vp, _i, _ll, _d = C.c_void_p, C.c_int, C.c_longlong, C.c_double

# name -> (restype, argtypes); must list every symbol of include/prysm_b200.h
SIGNATURES = {
    'pb_create': (_i, [C.POINTER(_vp), _i]),
    'pb_destroy': (_i, [_vp]),
    'pb_last_error': (C.c_char_p, [_vp]),
    'pb_version': (C.c_char_p, []),
    'pb_launch_count': (_ll, [_vp]),
    'pb_fft2': (_i, [_vp, _i, _vp, _i, _vp, _i, _d, _i, _i, _ll, _i, _i, _i, _d, _i, _i,
                     _vp, _i, _d, _i, _i, _ll, _vp]),
    'pb_fft2_batch': (_i, [_vp, _i, _vp, _i, _vp, _i, _d, _i, _ll, _ll, _i, _i, _ll, _i, _i, _i, _d, _i, _i,
                           _vp, _i, _d, _i, _i, _ll, _ll, _vp]),
    'pb_fft1': (_i, [_vp, _i, _vp, _i, _i, _ll, _i, _i, _i, _d, _vp, _ll, _vp]),
    'pb_axis_dft': (_i, [_vp, _i, _vp, _i, _i, _ll, _i, _i, _i, _d, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _ll, _vp]),
    'pb_czt_plan': (_i, [_vp, _i, _i, _i, _i, _d, _d, _i, _d, _d, _d, _vp, _vp, _vp, _vp, _vp]),
    'pb_czt_axis': (_i, [_vp, _i, _vp, _i, _i, _ll, _i, _i, _vp, _i, _vp, _vp, _i, _i, _i, _d, _vp, _ll, _vp]),
    'pb_czt_axis_intensity': (_i, [_vp, _i, _vp, _i, _i, _ll, _i, _i, _vp, _i, _vp, _vp, _i, _i, _i, _d, _i, _d, _vp, _ll, _vp]),
    'pb_polychromatic_czt_work_bytes': (_ll, [_i, _i, _i, _i]),
    'pb_polychromatic_czt': (_i, [_vp, _i, _vp, _i, _vp, _i, _i, _i, _i, C.POINTER(_d), _vp, _vp, _vp]),
    'pb_angular_spectrum': (_i, [_vp, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    'pb_angular_spectrum_screen': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _i, _vp]),
    'pb_angular_spectrum_vectors': (_i, [_vp, _i, _i, _i, _d, _d, _d, _vp, _vp, _vp]),
    'pb_mdft_basis': (_i, [_vp, _i, C.POINTER(_d), _i, C.POINTER(_d), _i, _i, _vp, _vp]),
    'pb_cgemm': (_i, [_vp, _i, _i, _i, _i, _i, _i, _d, _vp, _ll, _vp, _ll, _vp, _ll, _vp]),
    'pb_mdft_apply': (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _d, _i, _i, _vp, _vp]),
    'pb_mdft_work_elems': (_ll, [_i, _i, _i, _i, _i, _i]),
    'pb_mdft_tc_supported': (_i, [_i, _i, _i, _i]),
    'pb_mdft_tc_expand': (_i, [_vp, _vp, _i, _i, _vp, _vp, _vp]),
    'pb_mdft_tc_work_bytes': (_ll, [_i, _i, _i, _i]),
    'pb_mdft_tc_apply': (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _d, _vp, _vp]),
    'pb_phase_screen': (_i, [_vp, _i, _vp, _i, _vp, _d, _ll, _vp, _vp]),
    'pb_intensity': (_i, [_vp, _i, _vp, _ll, _d, _i, _vp, _vp]),
    'pb_binary': (_i, [_vp, _i, _i, _vp, _vp, _d, _d, _i, _ll, _vp, _vp]),
    'pb_mul_outer': (_i, [_vp, _i, _vp, _ll, _i, _i, _vp, _i, _vp, _i, _d, _vp, _ll, _vp]),
    'pb_weighted_sum': (_i, [_vp, _i, _vp, _i, _ll, C.POINTER(_d), _vp, _vp]),
    'pb_otf_normalize': (_i, [_vp, _i, _vp, _i, _i, _i, _vp, _vp, _vp, _vp]),
    'pb_encircled_energy': (_i, [_vp, _i, _vp, _i, _i, _d, C.POINTER(_d), _i, C.POINTER(_d), _vp]),
    'pb_moments': (_i, [_vp, _i, _vp, _i, _i, C.POINTER(_d), _vp]),
    'pb_mask_multiply': (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _vp, _d, _ll, _vp, _vp]),
    'pb_field_adjoint': (_i, [_vp, _i, _i, _vp, _vp, _vp, _d, _ll, _vp, _vp]),
    'pb_component': (_i, [_vp, _i, _i, _vp, _ll, _vp, _vp]),
    'pb_dot': (_i, [_vp, _i, _vp, _vp, _vp, _ll, C.POINTER(_d), _vp]),
    'pb_mode_projection': (_i, [_vp, _i, _vp, _i, _ll, _vp, _i, C.POINTER(_d), _vp]),
    'pb_otf_adjoint_seed': (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _vp, _vp]),
    'pb_encircled_energy_adjoint_seed': (_i, [_vp, _i, _i, _i, _d, C.POINTER(_d), C.POINTER(_d), _i, _vp, _vp]),
    'pb_vortex_phase': (_i, [_vp, _i, _i, _vp, _vp, _ll, _vp, _vp]),
    'pb_xy_grid': (_i, [_vp, _i, _i, _i, _d, _vp, _vp, _vp, _vp, _vp]),
    'pb_cart_to_polar': (_i, [_vp, _i, _vp, _vp, _ll, _vp, _vp, _vp]),
    'pb_circle': (_i, [_vp, _i, _vp, _ll, _d, _d, _vp, _vp]),
    'pb_jacobi_seq': (_i, [_vp, _i, _vp, _ll, _i, _d, _d, C.POINTER(_i), _vp, _vp]),
    'pb_zernike_seq': (_i, [_vp, _i, _i, _vp, _vp, _ll, _i, C.POINTER(_i), C.POINTER(_i), _i, _vp, _vp]),
    'pb_zernike_sum': (_i, [_vp, _i, _i, _vp, _vp, _ll, _i, C.POINTER(_i), C.POINTER(_i), C.POINTER(_d), _i, _vp, _vp]),
    'pb_balance_scale': (_i, [_vp, _i, _vp, _vp, _ll, _vp, _vp]),
    'pb_pack_complex': (_i, [_vp, _i, _vp, _vp, _vp, _ll, _vp, _vp]),
    'pb_packed_spectrum_product': (_i, [_vp, _i, _vp, _i, _i, _d, _vp, _vp, _vp]),
    'pb_resample_bilinear': (_i, [_vp, _i, _vp, _i, _i, _vp, _vp, _ll, _d, _d, _d, _vp, _d, _d, _vp, _vp]),
    'pb_bindown': (_i, [_vp, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    'pb_tile': (_i, [_vp, _i, _vp, _i, _i, _i, _i, _d, _vp, _vp]),
    'pb_separable_tf': (_i, [_vp, _i, _i, _vp, _vp, _i, _i, _d, _d, _vp, _vp]),
    'pb_radial_window': (_i, [_vp, _i, _i, _i, _d, _d, _d, _d, _d, _d, _vp, _vp, _vp, _vp]),
}


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f'{path} is missing: build it with `python prysm_b200/build.py` '
            '(prysm_b200 has no CPU fallback)')
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so and the header diverge
        fn.restype = res
        fn.argtypes = args
    return lib


lib = load_library()


class B200Error(RuntimeError):
    """A libprysm_b200 call failed (CUDA error, unsupported shape, bad argument)."""


class Handle:
    """One engine handle per CUDA device (twiddle tables + scratch live here)."""

    def __init__(self, device):
        self._h = _vp()
        rc = lib.pb_create(C.byref(self._h), int(device))
        if rc != 0:
            raise B200Error(f'pb_create(device={device}) failed with status {rc}: is a CUDA device visible?')
        self.device = int(device)

    def check(self, rc):
        if rc != 0:
            msg = lib.pb_last_error(self._h).decode()
            if rc == -1:
                raise ValueError(msg)
            raise B200Error(f'status {rc}: {msg}')

    @property
    def ptr(self):
        return self._h

    def launch_count(self):
        return int(lib.pb_launch_count(self._h))

    def close(self):
        if self._h:
            lib.pb_destroy(self._h)
            self._h = _vp()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_handles = {}          # (device, raw stream pointer) -> Handle, in least-recently-used order
_pinned = set()        # keys whose handle must outlive the cache (a captured CUDA graph replays into its scratch)
MAX_HANDLES_PER_DEVICE = 8


def handle_for(device_index, stream=0):
    """One handle per (device, stream): a handle's scratch arenas are only ordered by the stream its calls run
    on, so work issued on different streams must not share one (include/prysm_b200.h: 'not thread-safe').
    The cache is bounded: beyond MAX_HANDLES_PER_DEVICE the least recently used un-pinned handle of that device is
    destroyed (pb_destroy synchronises the device first), so short-lived side streams do not accumulate twiddle
    tables and scratch arenas that torch's allocator cannot see."""
    key = (device_index, stream)
    h = _handles.pop(key, None)
    if h is None:
        mine = [k for k in _handles if k[0] == device_index and k not in _pinned]
        while len(mine) >= MAX_HANDLES_PER_DEVICE:
            _handles.pop(mine.pop(0)).close()
        h = Handle(device_index)
    _handles[key] = h          # (re)insert as most recently used
    return h


def pin_handle(device_index, stream):
    """Keep the handle of (device, stream) alive for the life of the process (graphs.capture)."""
    _pinned.add((device_index, stream))


def release_handle(device_index, stream):
    """Destroy the handle of (device, stream) now (explicit close of a side stream's engine state)."""
    _pinned.discard((device_index, stream))
    h = _handles.pop((device_index, stream), None)
    if h is not None:
        h.close()


def launch_count(device_index):
    """Kernels launched on `device_index` through any stream's handle."""
    return sum(h.launch_count() for (d, _), h in _handles.items() if d == device_index)


def version():
    return lib.pb_version().decode()
