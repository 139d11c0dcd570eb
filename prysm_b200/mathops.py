"""Backend selection in the style of prysm.mathops (reference prysm/mathops.py:48-116).

`set_backend_to_b200()` is the sibling of `set_backend_to_cupy()`: it plugs this engine into an installed
prysm so that existing user code (`Wavefront.focus`, `free_space`, `prepare_executor`, `mtf_from_psf`, ...)
runs on the B200 kernels unchanged.  The reference's shim forwards *module attributes* (np.exp, fft.fft2);
fusing pad + shifts + FFT into one kernel needs the *function* level, so the hot functions themselves are
re-bound, following the precedent of prysm/x/polarization.py:541-552 (`setattr(propagation, name, wrapper)`).
`Wavefront` methods resolve `focus`, `angular_spectrum`, ... through the names imported into
prysm/propagation/wavefront.py:11-25, so those module globals are re-bound as well.

`set_backend_to_defaults()` restores every name it replaced.
"""
import importlib

from . import convolution as _conv
from . import fttools as _ft
from . import otf as _otf
from . import polynomials as _poly
from . import propagation as _prop
from . import psf as _psf

# prysm module -> {attribute: replacement}
_PROPAGATION_FUNCS = (
    'focus', 'unfocus', 'focus_adjoint', 'unfocus_adjoint',
    'angular_spectrum', 'angular_spectrum_adjoint', 'angular_spectrum_transfer_function',
    'prepare_executor', 'coordinates_for_focus', 'focus_dft', 'unfocus_dft', 'focus_dft_adjoint', 'unfocus_dft_adjoint',
)
_CORONAGRAPH_FUNCS = (
    'to_fpm_and_back', 'to_fpm_and_back_adjoint', 'to_fpm_and_back_multiresolution',
    'to_fpm_and_back_multiresolution_adjoint', 'babinet', 'babinet_adjoint', 'vortex_phase_mask', 'prepare_measured_fpm',
)
_TARGETS = {
    'prysm.propagation.fft': {n: getattr(_prop, n) for n in ('focus', 'unfocus', 'focus_adjoint', 'unfocus_adjoint')},
    'prysm.propagation.angular_spectrum': {n: getattr(_prop, n) for n in (
        'angular_spectrum', 'angular_spectrum_adjoint', 'angular_spectrum_transfer_function')},
    'prysm.propagation.dft': {**{n: getattr(_prop, n) for n in (
        'prepare_executor', 'coordinates_for_focus', 'focus_dft', 'unfocus_dft', 'focus_dft_adjoint', 'unfocus_dft_adjoint')},
        'MDFT': _ft.MDFT, 'CZT': _ft.CZT, 'FFTDFT': _ft.FFTDFT,
        'prepare_multiresolution': _prop.prepare_multiresolution, 'MultiResolutionExecutor': _prop.MultiResolutionExecutor},
    'prysm.propagation.coronagraph': {n: getattr(_prop, n) for n in _CORONAGRAPH_FUNCS},
    'prysm.propagation.wavefront': {**{n: getattr(_prop, n) for n in _PROPAGATION_FUNCS + _CORONAGRAPH_FUNCS if n not in (
        'angular_spectrum_transfer_function', 'coordinates_for_focus', 'vortex_phase_mask', 'prepare_measured_fpm')},
        'prepare_multiresolution': _prop.prepare_multiresolution, 'pad2d': _ft.pad2d, 'crop_center': _ft.crop_center},
    'prysm.propagation': {**{n: getattr(_prop, n) for n in _PROPAGATION_FUNCS + _CORONAGRAPH_FUNCS},
                          'prepare_multiresolution': _prop.prepare_multiresolution,
                          'MultiResolutionExecutor': _prop.MultiResolutionExecutor},
    'prysm.fttools': {'MDFT': _ft.MDFT, 'CZT': _ft.CZT, 'FFTDFT': _ft.FFTDFT, 'pad2d': _ft.pad2d,
                      'crop_center': _ft.crop_center, 'fourier_resample': _ft.fourier_resample},
    'prysm.convolution': {'conv': _conv.conv, 'apply_transfer_functions': _conv.apply_transfer_functions},
    'prysm.otf': {n: getattr(_otf, n) for n in (
        'transform_psf', 'transform_psf_adjoint', 'mtf_from_psf', 'ptf_from_psf', 'otf_from_psf', 'mtf_ptf_otf_from_psf',
        'mtf_from_psf_adjoint', 'ptf_from_psf_adjoint', 'otf_from_psf_adjoint', 'encircled_energy',
        'encircled_energy_adjoint')},
    'prysm.psf': {'centroid': _psf.centroid},
    'prysm.polynomials': {'sum_of_2d_modes': _poly.sum_of_2d_modes, 'sum_of_2d_modes_adjoint': _poly.sum_of_2d_modes_adjoint},
    'prysm.polynomials.fitting': {'sum_of_2d_modes': _poly.sum_of_2d_modes,
                                  'sum_of_2d_modes_adjoint': _poly.sum_of_2d_modes_adjoint},
}

_saved = {}  # (module name, attribute) -> original


def set_backend_to_b200(device=None):
    """Route prysm's propagation hot path through the B200 engine.  Call after `import prysm`
    (like the reference's own set_backend_to_* helpers).  Arrays handed to the re-bound functions may be
    numpy arrays (uploaded) or CUDA tensors; results are CUDA tensors (`prysm.mathops.array_to_true_numpy`
    already understands them, prysm/mathops.py:150-153)."""
    from . import _ops
    if device is not None:
        _ops.set_device(device)
    try:
        import prysm.conf as pconf
    except ImportError as exc:  # pragma: no cover
        raise ImportError('set_backend_to_b200() plugs into an installed prysm; it is not importable here') from exc
    from .conf import config
    config.precision = pconf.config.precision  # follow prysm's precision setting
    for modname, repl in _TARGETS.items():
        try:
            mod = importlib.import_module(modname)
        except ImportError:
            continue
        for attr, new in repl.items():
            if hasattr(mod, attr) and (modname, attr) not in _saved:
                _saved[(modname, attr)] = getattr(mod, attr)
            if hasattr(mod, attr):
                setattr(mod, attr, new)
    return sorted(_saved)


def set_backend_to_defaults():
    """Undo set_backend_to_b200(): every re-bound name gets its original object back."""
    for (modname, attr), orig in list(_saved.items()):
        setattr(importlib.import_module(modname), attr, orig)
        del _saved[(modname, attr)]
