"""``correlation`` with the pyccl signature (pyccl/correlations.py:45-127): C_ell -> xi(theta).

The call site ``lib.correlation_vec(cosmo, ell, C_ell, theta, type, method, len(theta), status)``
(pyccl/correlations.py:118-121) becomes one C-ABI call, ``ccl_b200_correlation`` (include/ccl_b200.h).  The 3-D
correlation functions of the same reference module (xi(r), multipoles, pi-sigma) are a different path
(src/ccl_correlation.c:442-974 over P(k)) and are not built.
"""
from enum import Enum

import numpy as np

from . import _lib, lowlevel
from .errors import check

__all__ = ("CorrelationMethods", "CorrelationTypes", "correlation")


class CorrelationMethods(Enum):
    """'Bessel' direct integration, 'FFTLog' fast Hankel transform, 'Legendre' brute-force sum."""
    FFTLOG = "fftlog"
    BESSEL = "bessel"
    LEGENDRE = "legendre"


class CorrelationTypes(Enum):
    NN = "NN"
    NG = "NG"
    GG_PLUS = "GG+"
    GG_MINUS = "GG-"


correlation_methods = {'fftlog': _lib.CORR_FFTLOG, 'bessel': _lib.CORR_BESSEL, 'legendre': _lib.CORR_LGNDRE}
correlation_types = {'NN': _lib.CORR_GG, 'NG': _lib.CORR_GL, 'GG+': _lib.CORR_LP, 'GG-': _lib.CORR_LM}


def correlation(cosmo, *, ell, C_ell, theta, type='NN', method='fftlog'):
    """Angular correlation function at ``theta`` (degrees) from ``C_ell`` sampled at ``ell``."""
    method = method.lower()
    if type not in correlation_types:
        raise ValueError(f"Invalid correlation type {type}.")
    if method not in correlation_methods.keys():
        raise ValueError(f"Invalid correlation method {method}.")
    if scalar := isinstance(theta, (int, float)):
        theta = np.array([theta, ])
    status = 0
    if np.all(np.array(C_ell) == 0):
        wth = np.zeros_like(theta)  # short-cut and also avoid integration errors
    else:
        if np.shape(ell) != np.shape(C_ell):
            from .errors import CCLError
            raise CCLError("Input shape for `larr` must match `clarr`!")  # pyccl/ccl_correlation.i:22-24
        wth, status = lowlevel.correlation(cosmo._device(), ell, C_ell, theta, correlation_types[type],
                                           correlation_methods[method])
    check(status, cosmo)
    if scalar:
        return wth[0]
    return wth
