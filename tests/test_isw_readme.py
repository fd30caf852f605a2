"""Two more of the reference's own checks for the path, on the pyccl-compatible API:
  * benchmarks/test_isw.py:6-49   ISW x number counts against a direct Simpson Limber integral (1e-3)
  * README.md:70-95               the README example (BBKS, two WeakLensingTracers with uniform n(z),
                                  ell = 2..2000, 'spline'): BASELINE.json configs[0]
The CPU variants run the oracle on the API's host-built tables; the GPU variants call angular_cl."""
import numpy as np
import pytest
from scipy.integrate import simpson

import oracle
from common import edge_mask, parity_error
from golden_setup import oracle_objects


def _isw_setup():
    import ccl_b200 as ccl
    Ob, Oc, h = 0.05, 0.25, 0.7
    cosmo = ccl.Cosmology(Omega_b=Ob, Omega_c=Oc, h=h, n_s=0.96, sigma8=0.8, transfer_function='bbks')
    ls = np.arange(2, 100).astype(float)
    zs = np.linspace(0, 0.6, 256)
    nz = np.exp(-0.5 * ((zs - 0.3) / 0.05) ** 2)
    tr_n = ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(zs, nz), bias=(zs, np.ones_like(zs)))
    tr_i = ccl.ISWTracer(cosmo)
    return cosmo, ls, zs, nz, tr_n, tr_i, (Ob, Oc, h)


def _isw_benchmark(cosmo, ls, zs, nz, pars, pk_eval):
    """Eq. 6 of arXiv:1710.03238 by direct Simpson integration (benchmarks/test_isw.py:30-46)."""
    import ccl_b200 as ccl
    Ob, Oc, h = pars
    pz = nz / simpson(nz, x=zs)
    H0 = h / ccl.physical_constants.CLIGHT_HMPC
    prefac = 3 * cosmo['T_CMB'] * (Oc + Ob) * H0 ** 3 / (ls + 0.5) ** 2
    ez = cosmo.h_over_h0(1. / (1 + zs))
    dz = cosmo.growth_factor(1. / (1 + zs))
    gz = np.gradient(dz * (1 + zs), zs[1] - zs[0]) / dz
    chi = cosmo.comoving_radial_distance(1 / (1 + zs))
    pks = np.array([pk_eval((ls + 0.5) / (c + 1E-6), 1. / (1 + z)) for c, z in zip(chi, zs)]).T
    return simpson(pks * (pz * ez * gz)[None, :], x=zs) * prefac


def test_isw_cpu():
    cosmo, ls, zs, nz, tr_n, tr_i, pars = _isw_setup()
    oc, op, otr = oracle_objects(cosmo, {"n": tr_n, "i": tr_i})
    cl, st = oracle.angular_cl(oc, otr["n"], otr["i"], op, ls, oracle.INTEG_QAG)
    assert st == 0
    clbb = _isw_benchmark(cosmo, ls, zs, nz, pars,
                          lambda k, a: np.array([op(np.log(kk), a, oc)[0] for kk in k]))
    assert np.all(np.fabs(cl / clbb - 1) < 1E-3)


@pytest.mark.gpu
def test_isw_gpu():
    import ccl_b200 as ccl
    cosmo, ls, zs, nz, tr_n, tr_i, pars = _isw_setup()
    cl = ccl.angular_cl(cosmo, tr_n, tr_i, ls)            # default method: qag_quad
    clbb = _isw_benchmark(cosmo, ls, zs, nz, pars, lambda k, a: ccl.nonlin_matter_power(cosmo, k, a))
    assert np.all(np.fabs(cl / clbb - 1) < 1E-3)
    oc, op, otr = oracle_objects(cosmo, {"n": tr_n, "i": tr_i})
    ref, st = oracle.angular_cl(oc, otr["n"], otr["i"], op, ls, oracle.INTEG_QAG)
    assert st == 0 and parity_error(cl, ref, None, 1e-5) < 1e-5


def _readme_setup():
    import ccl_b200 as ccl
    cosmo = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks')
    z = np.linspace(0., 1., 500)
    n = np.ones(z.shape)
    lens1 = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
    lens2 = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
    return cosmo, lens1, lens2, np.arange(2, 2001).astype(float)


def test_readme_example_cpu():
    cosmo, lens1, lens2, ell = _readme_setup()
    oc, op, otr = oracle_objects(cosmo, {"a": lens1, "b": lens2})
    ell = ell[::40]
    cl, st = oracle.angular_cl(oc, otr["a"], otr["b"], op, ell, oracle.INTEG_SPLINE)
    clq, stq = oracle.angular_cl(oc, otr["a"], otr["b"], op, ell, oracle.INTEG_QAG)
    assert st == 0 and stq == 0 and np.all(cl > 0) and np.all(np.diff(cl * ell * (ell + 1)) != 0)
    assert np.max(np.abs(cl / clq - 1)) < 1e-3          # the two Limber quadratures agree


@pytest.mark.gpu
def test_readme_example_gpu():
    import ccl_b200 as ccl
    cosmo, lens1, lens2, ell = _readme_setup()
    cls = ccl.angular_cl(cosmo, lens1, lens2, ell, limber_integration_method='spline')
    assert cls.shape == (1999,) and np.all(np.isfinite(cls)) and np.all(cls > 0)
    oc, op, otr = oracle_objects(cosmo, {"a": lens1, "b": lens2})
    redo = lambda: oracle.angular_cl(oc, otr["a"], otr["b"], op, ell, oracle.INTEG_SPLINE)[0]
    assert parity_error(cls, redo(), redo, 1e-10, edge_mask(otr["a"], otr["b"], ell)) < 1e-10
    clq = ccl.angular_cl(cosmo, lens1, lens2, ell[::50])
    refq, st = oracle.angular_cl(oc, otr["a"], otr["b"], op, ell[::50], oracle.INTEG_QAG)
    assert st == 0 and parity_error(clq, refq, None, 1e-5) < 1e-5
