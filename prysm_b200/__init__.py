"""prysm_b200 -- Blackwell-native engine for the prysm propagation hot path.

Mirrors the reference's module layout for the path only:
    prysm_b200.propagation  <-> prysm.propagation   (focus/unfocus/angular_spectrum/executors/Wavefront)
    prysm_b200.fttools      <-> prysm.fttools       (pad2d/crop_center/MDFT/CZT/FFTDFT)
    prysm_b200.otf / .psf   <-> prysm.otf / prysm.psf (transform_psf, mtf/ptf/otf, centroid)
    prysm_b200.polynomials  <-> prysm.polynomials.sum_of_2d_modes
    prysm_b200.conf         <-> prysm.conf          (config.precision)
    prysm_b200.mathops      <-> prysm.mathops       (set_backend_to_b200 / set_backend_to_defaults)
    prysm_b200.coronagraph  <-> prysm.propagation.coronagraph (re-exported by .propagation like the reference)
    prysm_b200.coordinates / .geometry  <-> prysm.coordinates.make_xy_grid / cart_to_polar, prysm.geometry.circle / antialias
    prysm_b200.convolution  <-> prysm.convolution (conv, apply_transfer_functions)
    prysm_b200.detector     <-> prysm.detector (bindown, tile, pixel_ft, olpf_ft)
    prysm_b200.graphs       CUDA-graph capture of launch-bound compositions (no reference counterpart)
Every array operation is a call into libprysm_b200.so (hand-written sm_100a CUDA behind the C ABI
of include/prysm_b200.h).  Importing this package without the built library raises ImportError.
"""
from . import _capi  # noqa: F401  (fails loudly if the CUDA library is missing)
from .conf import config  # noqa: F401
from . import fttools, propagation, coronagraph, otf, psf, polynomials, polychromatic, mathops, graphs  # noqa: F401
from . import coordinates, geometry, convolution, detector  # noqa: F401
from ._capi import B200Error  # noqa: F401
from .propagation import Wavefront  # noqa: F401
from ._ops import asdevice, asnumpy, set_device  # noqa: F401

__version__ = '0.1.0'
