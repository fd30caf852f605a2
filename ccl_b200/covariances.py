"""Limber C_ell covariances with the pyccl signatures (pyccl/covariances.py:10-123, 255-376): the connected
non-Gaussian and super-sample terms, both one call of ``ccl_angular_cl_covariance`` (src/ccl_cls.c:490-612) ->
``ccl_b200_angular_cl_covariance`` on the device."""
import numpy as np

from . import lowlevel
from .cells import integ_types
from .errors import check

__all__ = ("angular_cl_cov_cNG", "angular_cl_cov_SSC")


def _cov(cosmo, tracer1, tracer2, tracer3, tracer4, ell, ell2, t_of_kk_a, integration_method, chi_exponent,
         prefactor, kernel_extra):
    if integration_method not in integ_types:
        raise ValueError("Unknown integration method %s." % integration_method)
    cosmo.compute_distances()
    dev = cosmo._device()
    t1 = tracer1._device(cosmo)
    t2 = t1 if tracer2 is tracer1 else tracer2._device(cosmo)
    t3 = t1 if tracer3 is None else tracer3._device(cosmo)
    t4 = t2 if tracer4 is None else tracer4._device(cosmo)
    ell1_use = np.atleast_1d(np.asarray(ell, dtype=float))
    if ell2 is None:
        ell2 = ell
    ell2_use = np.atleast_1d(np.asarray(ell2, dtype=float))
    cov, status = lowlevel.angular_cl_covariance(dev, t1, t2, t3, t4, t_of_kk_a._device(), ell1_use, ell2_use,
                                                 integ_types[integration_method], chi_exponent, kernel_extra, prefactor)
    if np.ndim(ell2) == 0:
        cov = np.squeeze(cov, axis=0)
    if np.ndim(ell) == 0:
        cov = np.squeeze(cov, axis=-1)
    check(status, cosmo)
    return cov


def angular_cl_cov_cNG(cosmo, tracer1, tracer2, *, ell, t_of_kk_a, tracer3=None, tracer4=None, ell2=None, fsky=1.,
                       integration_method='qag_quad'):
    """Cov_cNG(l1, l2) = 1/(4 pi fsky) int dchi/chi^6 D1 D2 D3 D4 T((l1+1/2)/chi, (l2+1/2)/chi, a(chi)).
    ``out[i2, i1] = Cov(ell2[i2], ell[i1])``."""
    return _cov(cosmo, tracer1, tracer2, tracer3, tracer4, ell, ell2, t_of_kk_a, integration_method, 6,
                1. / (4 * np.pi * fsky), None)


def angular_cl_cov_SSC(cosmo, tracer1, tracer2, *, ell, t_of_kk_a, tracer3=None, tracer4=None, ell2=None,
                       sigma2_B=None, fsky=1., integration_method='qag_quad'):
    """Cov_SSC(l1, l2) = int dchi/chi^4 D1 D2 D3 D4 T(...) sigma_B^2(a(chi)).  ``sigma2_B`` = (a, sigma2_B(a)) must be
    given: the disc-variance helper ``sigma2_B_disc`` (pyccl/covariances.py:125-178, an integral over the linear
    power spectrum) is outside the Limber path of this build."""
    if sigma2_B is None:
        raise NotImplementedError("pass sigma2_B=(a, sigma2_B(a)); sigma2_B_disc is outside the Limber hot path")
    a_arr, s2b_arr = (np.ascontiguousarray(x, dtype=float) for x in sigma2_B)
    return _cov(cosmo, tracer1, tracer2, tracer3, tracer4, ell, ell2, t_of_kk_a, integration_method, 4, 1.,
                (a_arr, s2b_arr))
