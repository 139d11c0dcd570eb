"""Precision configuration, mirroring prysm.conf (reference prysm/conf.py:28-96).

``config.precision`` is a numpy real scalar type (float32 / float64) exactly as in the
reference; ``config.precision_complex`` is derived from it.  The CUDA kernels dispatch on the
dtype of the arrays they are handed (numpy's "infection" rule), and grids / bases / chirps
created by this package are created at ``config.precision``.
"""
from numbers import Integral

import numpy as np
import torch


def _coerce_real_dtype(precision):
    """Same acceptance rule as reference prysm/conf.py:8-20."""
    if isinstance(precision, Integral) and not isinstance(precision, bool):
        precision = f'float{precision}'
    try:
        dtype = np.dtype(precision)
    except (TypeError, ValueError) as exc:
        raise ValueError('precision should be a real floating dtype.') from exc
    if dtype.kind != 'f':
        raise ValueError('precision should be a real floating dtype.')
    return dtype.type


class Config:
    """Global configuration; only `precision` matters on the hot path."""

    def __init__(self, precision=64, *, lw=3, zorder=3):
        self.precision = precision
        self.lw = lw
        self.zorder = zorder

    @property
    def precision(self):
        return self._precision

    @precision.setter
    def precision(self, precision):
        p = _coerce_real_dtype(precision)
        if p not in (np.float32, np.float64):
            raise ValueError('the B200 engine computes in float32 or float64 (complex64 / complex128)')
        self._precision = p
        self._precision_complex = np.result_type(p, 1j).type

    @property
    def precision_complex(self):
        return self._precision_complex

    # torch views of the same setting
    @property
    def real_dtype(self):
        return torch.float32 if self._precision is np.float32 else torch.float64

    @property
    def complex_dtype(self):
        return torch.complex64 if self._precision is np.float32 else torch.complex128


config = Config()
