"""Limber C_ell covariance (ccl_angular_cl_covariance, src/ccl_cls.c:377-612; ccl_f3d_t, src/ccl_f3d.c).

Pin: the reference's own known-answer test, pyccl/tests/test_covariances.py::test_cov_NG_sanity -- a power-law
trispectrum T = 1/(k1^alpha k2^beta) and top-hat tracers on chi in [100, 1000] give the closed form
  Cov(l1, l2) = prefac (l1+1/2)^-alpha (l2+1/2)^-beta (chimax^x - chimin^x)/x,  x = alpha + beta - chi_power + 1,
asserted there at 1e-5 for 'qag_quad' and 4e-2 for 'spline' (cNG: chi_power 6, prefac 1/4pi; SSC: chi_power 4,
sigma2_B = 1).  The oracle restatement is held to the same numbers on the CPU; on the GPU the CUDA path is held to
them AND to the oracle (spline <= 1e-10, qag <= 1e-5), for the bicubic and the factorizable ccl_f3d_t forms."""
import numpy as np
import pytest

import oracle

NCHI, CHIMIN, CHIMAX = 100, 100., 1000.
CASES = [(3., 3., 'SSC'), (3., 2., 'SSC'), (2., 3., 'SSC'), (2., 2., 'SSC'),
         (3., 3., 'cNG'), (2., 2., 'cNG'), (1., 2., 'cNG'), (2., 1., 'cNG')]
LS = np.array([2., 20., 200.])


def _arrays(alpha, beta):
    a_arr = np.linspace(0.1, 1., 10)
    k_arr = np.geomspace(1E-4, 1E3, 10)
    tkka = np.array([1. / (k_arr[None, :] ** alpha * k_arr[:, None] ** beta) for _ in a_arr])
    return a_arr, k_arr, tkka


def pred_covar(l1, l2, alpha=1., beta=1., prefac=1. / (4 * np.pi), chi_power=6):
    lfac = 1 / ((l1 + 0.5) ** alpha * (l2 + 0.5) ** beta)
    xpn = alpha + beta - chi_power + 1
    return lfac * (CHIMAX ** xpn - CHIMIN ** xpn) / xpn * prefac


def _kw(typ):
    if typ == 'cNG':
        return dict(chi_exponent=6, prefactor=1. / (4 * np.pi), kernel_extra=None), dict()
    a_s = np.linspace(0.1, 1., 1024)
    return dict(chi_exponent=4, prefactor=1., kernel_extra=(a_s, np.ones_like(a_s))), dict(prefac=1., chi_power=4)


@pytest.fixture(scope="module")
def cosmo():
    import ccl_b200 as ccl
    c = ccl.CosmologyVanillaLCDM(transfer_function='bbks')
    c.compute_distances()
    return c


def _oracle_side(cosmo, alpha, beta):
    a_arr, k_arr, tkka = _arrays(alpha, beta)
    b = cosmo._bg
    oc = oracle.Cosmo(b.achi_chi, b.achi_a, None, None)
    tsp = oracle.F3D(a_arr, np.log(k_arr), tkka=np.log(tkka), is_log=True)
    chis = np.linspace(CHIMIN, CHIMAX, NCHI)
    tr = [oracle.Tracer(kernel=(chis, np.ones(NCHI)), chi_max_nokernel=cosmo.chi_max_nokernel())]
    return oc, tsp, tr


@pytest.mark.parametrize("alpha,beta,typ", CASES)
def test_oracle_covariance_closed_form(cosmo, alpha, beta, typ):
    oc, tsp, tr = _oracle_side(cosmo, alpha, beta)
    kw, pk = _kw(typ)
    cov_p = pred_covar(LS[None, :], LS[:, None], alpha, beta, **pk)
    cq, st = oracle.angular_cl_cov(oc, tr, tr, tr, tr, tsp, LS, LS, oracle.INTEG_QAG, **kw)
    assert st == 0 and np.all(np.fabs(cq / cov_p - 1) < 1E-5)                  # the reference's assertion
    cs, st = oracle.angular_cl_cov(oc, tr, tr, tr, tr, tsp, LS, LS, oracle.INTEG_SPLINE, **kw)
    assert st == 0 and np.all(np.fabs(cs / cov_p - 1) < 4E-2)


def test_oracle_f3d_forms_agree(cosmo):
    """A separable trispectrum through the bicubic form and through the factorizable form (two ccl_f2d_t)."""
    a_arr, k_arr, tkka = _arrays(2., 1.)
    lk = np.log(k_arr)
    full = oracle.F3D(a_arr, lk, tkka=np.log(tkka), is_log=True)
    p1 = np.array([-2. * lk for _ in a_arr])       # log of k^-2 and k^-1: is_log applies to each factor
    p2 = np.array([-1. * lk for _ in a_arr])
    prod = oracle.F3D(a_arr, lk, fka1=p1, fka2=p2, is_log=True)
    for (x, y, a) in ((-3.0, 1.0, 0.5), (0.3, -6.0, 0.95), (8.0, -12.0, 0.05), (2.0, 2.0, 1.0)):   # incl. extrapolation
        v1, s1 = full(x, y, a)
        v2, s2 = prod(x, y, a)
        assert s1 == 0 and s2 == 0 and abs(v1 / v2 - 1) < 1e-9, (x, y, a, v1, v2)


@pytest.mark.gpu
@pytest.mark.parametrize("alpha,beta,typ", CASES)
def test_gpu_covariance_closed_form_and_oracle(cosmo, alpha, beta, typ):
    """The reference's test_cov_NG_sanity through the pyccl-compatible API, plus parity with the oracle."""
    import ccl_b200 as ccl
    a_arr, k_arr, tkka = _arrays(alpha, beta)
    tsp = ccl.Tk3D(a_arr=a_arr, lk_arr=np.log(k_arr), tkk_arr=np.log(tkka), is_logt=True)
    tr = ccl.Tracer()
    tr.add_tracer(cosmo, kernel=(np.linspace(CHIMIN, CHIMAX, NCHI), np.ones(NCHI)))
    kw, pk = _kw(typ)
    cov_p = pred_covar(LS[None, :], LS[:, None], alpha, beta, **pk)
    if typ == 'cNG':
        def cov_f(ll, **kwargs):
            return ccl.angular_cl_cov_cNG(cosmo, tr, tr, ell=ll, t_of_kk_a=tsp, **kwargs)
    else:
        def cov_f(ll, **kwargs):
            return ccl.angular_cl_cov_SSC(cosmo, tr, tr, ell=ll, t_of_kk_a=tsp, sigma2_B=kw["kernel_extra"], **kwargs)
    cov = cov_f(LS)
    assert np.all(np.fabs(cov / cov_p - 1).flatten() < 1E-5)
    cov_s = cov_f(LS, integration_method='spline')
    assert np.all(np.fabs(cov_s / cov_p - 1).flatten() < 4E-2)
    assert np.all(np.fabs(cov_f(LS, tracer3=tr, tracer4=tr) / cov_p - 1).flatten() < 1E-5)
    assert np.all(np.fabs(np.array([cov_f(LS, ell2=l) for l in LS]) / cov_p - 1).flatten() < 1E-5)
    assert np.all(np.fabs(np.array([cov_f(l, ell2=LS) for l in LS]).T / cov_p - 1).flatten() < 1E-5)
    assert np.all(np.fabs(np.array([[cov_f(l1, ell2=l2) for l1 in LS] for l2 in LS]) / cov_p - 1).flatten() < 1E-5)
    # parity with the oracle on the same tables
    oc, otsp, otr = _oracle_side(cosmo, alpha, beta)
    rq, st = oracle.angular_cl_cov(oc, otr, otr, otr, otr, otsp, LS, LS, oracle.INTEG_QAG, **kw)
    rs, st2 = oracle.angular_cl_cov(oc, otr, otr, otr, otr, otsp, LS, LS, oracle.INTEG_SPLINE, **kw)
    assert st == 0 and st2 == 0
    assert np.max(np.abs(cov / rq - 1)) < 1e-5
    assert np.max(np.abs(cov_s / rs - 1)) < 1e-10


@pytest.mark.gpu
def test_gpu_covariance_realistic_tracers_and_product_form(cosmo):
    """Lensing x clustering tracers, a factorizable Tk3D with a-dependence, distinct ell sets: CUDA vs oracle."""
    import ccl_b200 as ccl
    from golden_setup import oracle_objects
    z = np.linspace(0., 1.2, 200)
    n = np.exp(-((z - 0.6) / 0.2) ** 2)
    wl = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
    nc = ccl.NumberCountsTracer(cosmo, has_rsd=True, dndz=(z, n), bias=(z, 1 + z))   # the RSD term drops out
    a_arr = np.linspace(0.3, 1., 12)
    lk = np.log(np.geomspace(1e-4, 50., 40))
    p1 = np.array([np.log(a ** 2 * 1e4 / (1 + (np.exp(lk) / 0.02) ** 2.5)) for a in a_arr])
    p2 = np.array([np.log(a * 3e3 / (1 + (np.exp(lk) / 0.05) ** 2.2)) for a in a_arr])
    tsp = ccl.Tk3D(a_arr=a_arr, lk_arr=lk, pk1_arr=p1, pk2_arr=p2, is_logt=True)
    otsp = oracle.F3D(a_arr, lk, fka1=p1, fka2=p2, is_log=True)
    oc, op, otr = oracle_objects(cosmo, {"wl": wl, "nc": nc})
    l1, l2 = np.array([10., 100., 800.]), np.array([30., 300.])
    for method, om, tol in (("qag_quad", oracle.INTEG_QAG, 1e-5), ("spline", oracle.INTEG_SPLINE, 1e-10)):
        cov = ccl.angular_cl_cov_cNG(cosmo, wl, nc, ell=l1, t_of_kk_a=tsp, tracer3=nc, tracer4=wl, ell2=l2, fsky=0.4,
                                     integration_method=method)
        ref, st = oracle.angular_cl_cov(oc, otr["wl"], otr["nc"], otr["nc"], otr["wl"], otsp, l1, l2, om, chi_exponent=6,
                                        prefactor=1. / (4 * np.pi * 0.4))
        assert st == 0 and cov.shape == (2, 3) and np.all(np.isfinite(cov)) and np.all(cov != 0)
        assert np.max(np.abs(cov / ref - 1)) < tol, method
    with pytest.raises(ValueError):
        ccl.angular_cl_cov_cNG(cosmo, wl, wl, ell=l1, t_of_kk_a=tsp, integration_method='cag_cuad')
