"""Backend selection in the style of prysm.mathops (reference prysm/mathops.py:48-116).

`set_backend_to_b200()` is the sibling of `set_backend_to_cupy()`: it plugs this engine into an installed
prysm so that existing user code (`Wavefront.focus`, `free_space`, `prepare_executor`, `mtf_from_psf`, ...)
runs on the B200 kernels unchanged.  The reference's shim forwards *module attributes* (np.exp, fft.fft2);
fusing pad + shifts + FFT into one kernel needs the *function* level, so the hot functions themselves are
re-bound, following the precedent of prysm/x/polarization.py:541-552 (`setattr(propagation, name, wrapper)`).
`Wavefront` methods resolve `focus`, `angular_spectrum`, ... through the names imported into
prysm/propagation/wavefront.py:11-25, so those module globals are re-bound as well.

The reference's `Wavefront` class itself is patched in place as well (its elementwise constructors and
properties are written against the `np` shim -- `np.exp` on the host, three temporaries for `|.|^2`,
prysm/propagation/wavefront.py:59-166): `from_amp_and_phase`, `phase_screen`, `thin_lens`, `intensity`, `phase`,
`real`, `imag`, the `* / + -` operators and the adjoint twins delegate to the kernels and return objects of the
reference's own classes, so user code that imported `Wavefront` before the switch keeps working and never
touches numpy `exp` or eager tensor arithmetic.

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
_saved_cls = {}  # (class, attribute) -> original descriptor (or _MISSING)
_MISSING = object()


def _wavefront_patches(ref_cls, ref_richdata):
    """Descriptors that replace the elementwise members of the reference's Wavefront class.  Each one builds the
    engine's Wavefront around the same device array, calls the engine, and hands the result back as an object of
    the REFERENCE's classes (so isinstance checks and attribute access in user code are unchanged)."""
    from . import _ops
    from ._richdata import RichData as _OurRich
    Ours = _prop.Wavefront

    def to_ours(w):
        return Ours(w.data, w.wavelength, w.dx, w.space)

    def unwrap(v):
        if isinstance(v, ref_cls):
            return to_ours(v)
        if isinstance(v, ref_richdata):
            return _OurRich(v.data, v.dx, v.wavelength)
        return v

    def rewrap(v):
        if isinstance(v, Ours):
            return ref_cls(v.data, v.wavelength, v.dx, v.space)
        if isinstance(v, _OurRich):
            return ref_richdata(v.data, v.dx, v.wavelength)
        if isinstance(v, tuple):
            return tuple(rewrap(e) for e in v)
        return v

    def method(name):
        def f(self, *args, **kwargs):
            r = getattr(to_ours(self), name)(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in kwargs.items()})
            return rewrap(r)
        f.__name__ = name
        f.__doc__ = getattr(Ours, name).__doc__
        return f

    def classm(name):
        def f(cls, *args, **kwargs):
            r = getattr(Ours, name)(*[unwrap(a) for a in args], **{k: unwrap(v) for k, v in kwargs.items()})
            if isinstance(r, Ours):
                return cls(r.data, r.wavelength, r.dx, r.space)
            return rewrap(r)
        f.__name__ = name
        f.__doc__ = getattr(Ours, name).__doc__
        return classmethod(f)

    def prop(name):
        return property(lambda self: rewrap(getattr(to_ours(self), name)), doc=getattr(Ours, name).__doc__)

    out = {n: classm(n) for n in ('from_amp_and_phase', 'phase_screen', 'thin_lens', 'thin_lens_adjoint')}
    out.update({n: prop(n) for n in ('intensity', 'phase', 'real', 'imag')})
    out.update({n: method(n) for n in ('__numerical_operation__', 'intensity_adjoint', 'from_amp_and_phase_adjoint_phase',
                                       'from_amp_and_phase_adjoint_amp', 'phase_screen_adjoint_phase')})
    return out


def _returning_reference_richdata(fn):
    """Wrap an engine function so that every engine RichData in its result becomes a prysm._richdata.RichData."""
    import functools
    from ._richdata import RichData as _OurRich
    ref = importlib.import_module('prysm._richdata').RichData

    def conv(v):
        if isinstance(v, _OurRich):
            return ref(v.data, v.dx, v.wavelength)
        if isinstance(v, tuple):
            return tuple(conv(e) for e in v)
        return v

    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        return conv(fn(*args, **kwargs))
    wrapped.__wrapped_engine__ = fn
    return wrapped


def _patch_wavefront_class():
    try:
        pw = importlib.import_module('prysm.propagation.wavefront')
        rd = importlib.import_module('prysm._richdata').RichData
    except ImportError:
        return
    cls = pw.Wavefront
    for name, desc in _wavefront_patches(cls, rd).items():
        if (cls, name) not in _saved_cls:
            _saved_cls[(cls, name)] = cls.__dict__.get(name, _MISSING)
        setattr(cls, name, desc)


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
        if modname == 'prysm.otf':   # these return RichData: hand back the reference's own container
            repl = {k: _returning_reference_richdata(v) for k, v in repl.items()}
        for attr, new in repl.items():
            if hasattr(mod, attr) and (modname, attr) not in _saved:
                _saved[(modname, attr)] = getattr(mod, attr)
            if hasattr(mod, attr):
                setattr(mod, attr, new)
    _patch_wavefront_class()
    return sorted(_saved)


def set_backend_to_defaults():
    """Undo set_backend_to_b200(): every re-bound name gets its original object back."""
    for (modname, attr), orig in list(_saved.items()):
        setattr(importlib.import_module(modname), attr, orig)
        del _saved[(modname, attr)]
    for (cls, name), orig in list(_saved_cls.items()):
        if orig is _MISSING:
            delattr(cls, name)
        else:
            setattr(cls, name, orig)
        del _saved_cls[(cls, name)]
