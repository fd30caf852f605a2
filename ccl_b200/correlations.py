"""``correlation`` with the pyccl signature (pyccl/correlations.py:45-127): C_ell -> xi(theta).

The call site ``lib.correlation_vec(cosmo, ell, C_ell, theta, type, method, len(theta), status)``
(pyccl/correlations.py:118-121) becomes one C-ABI call, ``ccl_b200_correlation`` (include/ccl_b200.h).  The 3-D
correlation functions of the same reference module (xi(r), multipoles, pi-sigma) are a different path
(src/ccl_correlation.c:442-974 over P(k)) and are not built.
"""
from enum import Enum

import numpy as np

from . import _lib, lowlevel
from .errors import CCLError, check

__all__ = ("CorrelationMethods", "CorrelationTypes", "correlation")

# name -> (user-facing string, C enum of include/ccl_correlation.h:6-12)
_METHODS = {"FFTLOG": ("fftlog", _lib.CORR_FFTLOG), "BESSEL": ("bessel", _lib.CORR_BESSEL),
            "LEGENDRE": ("legendre", _lib.CORR_LGNDRE)}
_TYPES = {"NN": ("NN", _lib.CORR_GG), "NG": ("NG", _lib.CORR_GL), "GG_PLUS": ("GG+", _lib.CORR_LP),
          "GG_MINUS": ("GG-", _lib.CORR_LM)}

#: 'bessel' direct integration, 'fftlog' fast Hankel transform, 'legendre' brute-force sum over multipoles
CorrelationMethods = Enum("CorrelationMethods", {k: v[0] for k, v in _METHODS.items()})
#: spin combinations: 0x0, 0x2, 2x2 (xi+), 2x2 (xi-)
CorrelationTypes = Enum("CorrelationTypes", {k: v[0] for k, v in _TYPES.items()})

correlation_methods = {label: code for label, code in _METHODS.values()}
correlation_types = {label: code for label, code in _TYPES.values()}


def correlation(cosmo, *, ell, C_ell, theta, type='NN', method='fftlog'):
    """Angular correlation function at ``theta`` (degrees) from ``C_ell`` sampled at ``ell``; ``type`` in
    'NN', 'NG', 'GG+', 'GG-' and ``method`` in 'fftlog', 'bessel', 'legendre' (case-insensitive)."""
    code_m = correlation_methods.get(method.lower())
    code_t = correlation_types.get(type)
    if code_t is None:
        raise ValueError(f"Invalid correlation type {type}.")
    if code_m is None:
        raise ValueError(f"Invalid correlation method {method.lower()}.")
    is_scalar = isinstance(theta, (int, float))
    th = np.array([theta]) if is_scalar else theta
    if not np.any(np.asarray(C_ell)):
        # an all-zero spectrum transforms to zero (and its power-law extrapolation has no slope to take)
        xi = np.zeros_like(th)
    else:
        if np.shape(ell) != np.shape(C_ell):
            raise CCLError("Input shape for `larr` must match `clarr`!")  # pyccl/ccl_correlation.i:22-24
        xi, status = lowlevel.correlation(cosmo._device(), ell, C_ell, th, code_t, code_m)
        check(status, cosmo)
    return xi[0] if is_scalar else xi
