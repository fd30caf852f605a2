"""C_ell -> xi(theta) (ccl_correlation, src/ccl_correlation.c:14-440; FFTLog, src/ccl_fftlog.c).

Pins of the oracle restatement (CPU):
  * the reference's angular-correlation benchmark (benchmarks/test_correlation.py:12-231): 28 tracer-pair cases x 15
    angles x 2 n(z) families from an independent code, asserted at the reference's own 0.2 ('fftlog') / 0.1
    ('bessel') of the LSST error bars (tests/golden/correlation_benchmark.npz, extracted by tools/make_golden.py);
  * closed-form Hankel pairs of a Gaussian C_ell (J0 and J2), the complex log-Gamma against scipy.
GPU: the CUDA path (ccl_b200_correlation through `ccl_b200.correlation`) against the oracle on the same C_ell for
every method and type (FFTLog / Legendre <= 1e-9 of the curve's scale, Bessel <= 1e-5: adaptive), against the
benchmark, and the reference's unit tests (pyccl/tests/test_correlations.py:13-33,105-135)."""
import os

import numpy as np
import pytest
from scipy.interpolate import interp1d

import oracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "correlation_benchmark.npz")

# (tracer1, tracer2, benchmark, error bar, type, prefactor): benchmarks/test_correlation.py:176-203
CASES = [('g1', 'g1', 'b1b1_dd', ('clustering', 1), 'NN', 1), ('g2', 'g2', 'b2b2_dd', ('clustering', 2), 'NN', 1),
         ('g1', 'l1', 'b1b1_dl', ('ggl', 2), 'NG', 1), ('g1', 'l2', 'b1b2_dl', ('ggl', 1), 'NG', 1),
         ('g2', 'l1', 'b2b1_dl', ('ggl', 4), 'NG', 1), ('g2', 'l2', 'b2b2_dl', ('ggl', 3), 'NG', 1),
         ('g1', 'i1', 'b1b1_di', ('ggl', 2), 'NG', 1), ('g1', 'i2', 'b1b2_di', ('ggl', 1), 'NG', 1),
         ('g2', 'i1', 'b2b1_di', ('ggl', 4), 'NG', 1), ('g2', 'i2', 'b2b2_di', ('ggl', 3), 'NG', 1),
         ('l1', 'l1', 'b1b1_ll_pp', ('xi+', 1), 'GG+', 1), ('l1', 'l2', 'b1b2_ll_pp', ('xi+', 3), 'GG+', 1),
         ('l2', 'l2', 'b2b2_ll_pp', ('xi+', 2), 'GG+', 1), ('l1', 'l1', 'b1b1_ll_mm', ('xi-', 1), 'GG-', 1),
         ('l1', 'l2', 'b1b2_ll_mm', ('xi-', 3), 'GG-', 1), ('l2', 'l2', 'b2b2_ll_mm', ('xi-', 2), 'GG-', 1),
         ('i1', 'l1', 'b1b1_li_pp', ('xi+', 1), 'GG+', 2), ('i1', 'l2', 'b1b2_li_pp', ('xi+', 1), 'GG+', 1),
         ('i2', 'l2', 'b2b2_li_pp', ('xi+', 2), 'GG+', 2), ('i1', 'l1', 'b1b1_li_mm', ('xi-', 1), 'GG-', 2),
         ('i1', 'l2', 'b1b2_li_mm', ('xi-', 1), 'GG-', 1), ('i2', 'l2', 'b2b2_li_mm', ('xi-', 2), 'GG-', 2),
         ('i1', 'i1', 'b1b1_ii_pp', ('xi+', 1), 'GG+', 1), ('i1', 'i2', 'b1b2_ii_pp', ('xi+', 3), 'GG+', 1),
         ('i2', 'i2', 'b2b2_ii_pp', ('xi+', 2), 'GG+', 1), ('i1', 'i1', 'b1b1_ii_mm', ('xi-', 1), 'GG-', 1),
         ('i1', 'i2', 'b1b2_ii_mm', ('xi-', 3), 'GG-', 1), ('i2', 'i2', 'b2b2_ii_mm', ('xi-', 2), 'GG-', 1)]
ERRFAC = {'fftlog': 0.2, 'bessel': 0.1}
EPSREL = 2.5E-5  # ccl.gsl_params.INTEGRATION_LIMBER_EPSREL / INTEGRATION_EPSREL of the benchmark (:25-26)


def bench_ells():
    """benchmarks/test_correlation.py:34-42"""
    lmax = 10000
    nls = (lmax - 400) // 20 + 141
    ells = np.zeros(nls)
    ells[:101] = np.arange(101)
    ells[101:121] = ells[100] + (np.arange(20) + 1) * 5
    ells[121:141] = ells[120] + (np.arange(20) + 1) * 10
    ells[141:] = ells[140] + (np.arange(nls - 141) + 1) * 20
    return lmax, ells


def error_bar(d, er, theta):
    """benchmarks/test_correlation.py:128-171"""
    tab = d["sigma_" + er[0]]
    return interp1d(tab[0] / 60., tab[er[1]], fill_value=tab[er[1]][0], bounds_error=False)(theta)


_cl_cache = {}


def benchmark_cl(nztyp, t1, t2):
    """C_ell of the benchmark's tracers from the oracle's 'qag_quad' path at the benchmark's tolerance,
    interpolated to every integer multipole like the reference test (:212-216)."""
    key = (nztyp, t1, t2)
    if key not in _cl_cache:
        from golden_setup import oracle_objects, set_up
        cosmo, trc, _, _ = set_up(nztyp)
        if nztyp not in _cl_cache:
            oc, op, otr = oracle_objects(cosmo, trc)
            bg, sp = cosmo._bg, cosmo._spline_params
            oc = oracle.Cosmo(bg.achi_chi, bg.achi_a, bg.a, bg.growth, K_MAX=sp["K_MAX"], K_MIN=sp["K_MIN"],
                              DLOGK=sp["DLOGK_INTEGRATION"], LIMBER_EPSREL=EPSREL, N_ITERATION=1000)
            _cl_cache[nztyp] = (oc, op, otr)
        oc, op, otr = _cl_cache[nztyp]
        lmax, ells = bench_ells()
        cl, st = oracle.angular_cl(oc, otr[t1], otr[t2], op, ells, oracle.INTEG_QAG)
        assert st == 0
        ell = np.arange(lmax)
        _cl_cache[key] = (ell.astype(float), interp1d(ells, cl, kind='cubic')(ell))
    return _cl_cache[key]


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


# ------------------------------------------------------------------ CPU: oracle pins
@pytest.mark.parametrize("nztyp", ["analytic", "histo"])
@pytest.mark.parametrize("method", ["fftlog", "bessel"])
def test_oracle_correlation_reference_benchmark(golden, nztyp, method):
    theta = golden["theta"]
    worst = 0.0
    for t1, t2, bm, er, kind, pref in CASES:
        ell, cli = benchmark_cl(nztyp, t1, t2)
        xi, st = oracle.correlation(ell, cli, theta, kind, method, epsrel=EPSREL)
        assert st == 0
        dev = np.fabs(xi * pref - golden["%s_%s" % (nztyp, bm)]) / (error_bar(golden, er, theta) * ERRFAC[method])
        worst = max(worst, dev.max())
        assert np.all(dev < 1), (t1, t2, kind, dev.max())
    assert worst < 1


def test_oracle_lngamma_complex_vs_scipy():
    from scipy.special import loggamma
    for zr in (0.5, 1.5, 2.5, 0.25, 3.0):
        for zi in (0.0, 0.3, 7.0, 61.0, 250.0, 503.0, -77.0):
            lnr, arg = oracle.lngamma_complex(zr, zi)
            ref = loggamma(complex(zr, zi))
            assert abs(lnr - ref.real) <= 1e-12 * max(1.0, abs(ref.real))
            assert abs(np.angle(np.exp(1j * (arg - ref.imag)))) <= 2e-12 * max(1.0, abs(ref.imag))
            assert -np.pi <= arg <= np.pi


def gaussian_cl():
    s = 0.003  # rad
    ell = np.geomspace(0.5, 3000, 300)
    return s, ell, np.exp(-0.5 * (ell * s) ** 2)


def test_oracle_correlation_closed_form():
    """xi(theta) = int l dl / (2 pi) C_l J_n(l theta) of a Gaussian C_l, n = 0 and 2"""
    s, ell, cl = gaussian_cl()
    th = np.array([0.05, 0.1, 0.2, 0.3])
    x = np.radians(th) / s
    ex0 = np.exp(-0.5 * x ** 2) / (2 * np.pi * s * s)
    ex2 = (2 / x ** 2 * (1 - np.exp(-x * x / 2)) - np.exp(-x * x / 2)) / (2 * np.pi * s * s)
    for kind, ex in (("NN", ex0), ("GG+", ex0), ("NG", ex2)):
        w, st = oracle.correlation(ell, cl, th, kind, "bessel")
        assert st == 0 and np.all(np.fabs(w / ex - 1) < 1e-4)       # QAG at epsrel 1e-4
        w, st = oracle.correlation(ell, cl, th, kind, "fftlog")
        assert st == 0 and np.all(np.fabs(w - ex) < 1.2e-2 * np.max(ex))  # FFTLog: ringing + power-law extrapolation
        if kind != "GG+":
            w, st = oracle.correlation(ell, cl, th, kind, "legendre")
            assert st == 0 and np.all(np.fabs(w - ex) < 5e-3 * np.max(ex))  # full sky vs flat sky
    w, st = oracle.correlation(ell, cl, th, "GG+", "legendre")
    assert st == 1040                                                  # CCL_ERROR_NOT_IMPLEMENTED (:308-313)


def test_oracle_correlation_errors():
    """src/ccl_f1d.c:97-102: a non-positive product of the last two C_ell makes the spline fail"""
    ell = np.arange(2., 1001.)
    cl = np.zeros(ell.size)
    cl[500] = 1.0
    for method, code in (("fftlog", 1028), ("bessel", 1025), ("legendre", 1025)):
        w, st = oracle.correlation(ell, cl, [1.0, 2.0], "NN", method)
        assert st == code
    w, st = oracle.correlation(ell[:4], np.ones(4), [1.0], "NN", "fftlog")  # akima needs 5 nodes
    assert st == 1025
    lim = [2., 10., 500., 900.]
    s, ell, cl = gaussian_cl()
    w0, st0 = oracle.correlation(ell, cl, [0.1, 0.3], "NN", "fftlog")
    w1, st1 = oracle.correlation(ell, cl, [0.1, 0.3], "NN", "fftlog", taper_cl_limits=lim)
    assert st0 == 0 and st1 == 0 and np.all(np.isfinite(w1)) and not np.allclose(w0, w1, rtol=1e-6)


# ------------------------------------------------------------------ GPU: CUDA path vs oracle
@pytest.fixture(scope="module")
def cosmo():
    import ccl_b200 as ccl
    return ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks',
                         matter_power_spectrum='linear')


def _scale(ref):
    return np.max(np.fabs(ref))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["NN", "NG", "GG+", "GG-"])
@pytest.mark.parametrize("method,tol", [("fftlog", 1e-9), ("legendre", 1e-9), ("bessel", 1e-5)])
def test_device_correlation_matches_oracle(cosmo, kind, method, tol):
    import ccl_b200 as ccl
    s, ell, cl = gaussian_cl()
    cl = cl * (1.0 + 0.3 * np.sin(np.log(ell) * 3.0)) + 1e-9 * (ell / 100.) ** -1.2
    theta = np.geomspace(0.02, 4.0, 23)
    ref, st = oracle.correlation(ell, cl, theta, kind, method)
    if method == "legendre" and kind in ("GG+", "GG-"):
        assert st == 1040
        with pytest.raises(ccl.CCLError):
            ccl.correlation(cosmo, ell=ell, C_ell=cl, theta=theta, type=kind, method=method)
        return
    assert st == 0
    xi = ccl.correlation(cosmo, ell=ell, C_ell=cl, theta=theta, type=kind, method=method)
    assert xi.shape == theta.shape and np.all(np.isfinite(xi))
    assert np.max(np.fabs(xi - ref)) <= tol * _scale(ref), np.max(np.fabs(xi - ref)) / _scale(ref)


@pytest.mark.gpu
@pytest.mark.parametrize("method", ["fftlog", "bessel"])
def test_device_correlation_reference_benchmark(golden, method):
    """The CUDA path on the reference's benchmark at the reference's tolerance (analytic n(z); C_ell from the
    device 'qag_quad' path through the API like the reference test)."""
    import ccl_b200 as ccl
    from golden_setup import set_up
    cosmo, trc, _, _ = set_up("analytic")
    ccl.gsl_params.INTEGRATION_EPSREL = EPSREL
    try:
        c2 = ccl.Cosmology(Omega_c=0.30, Omega_b=0.00, Omega_g=0, Omega_k=0, h=0.7, sigma8=0.8, n_s=0.96, Neff=0,
                           m_nu=0.0, w0=-1, wa=0, T_CMB=2.7, transfer_function='bbks', matter_power_spectrum='linear')
    finally:
        ccl.gsl_params.reload()
    theta = golden["theta"]
    lmax, ells = bench_ells()
    ell = np.arange(lmax)
    for t1, t2, bm, er, kind, pref in CASES[::3]:
        cl = ccl.angular_cl(cosmo, trc[t1], trc[t2], ells)
        cli = interp1d(ells, cl, kind='cubic')(ell)
        xi = ccl.correlation(c2, ell=ell, C_ell=cli, theta=theta, type=kind, method=method)
        dev = np.fabs(xi * pref - golden["analytic_" + bm]) / (error_bar(golden, er, theta) * ERRFAC[method])
        assert np.all(dev < 1), (t1, t2, kind, dev.max())
        ref, st = oracle.correlation(ell.astype(float), cli, theta, kind, method, epsrel=EPSREL)
        assert st == 0
        assert np.max(np.fabs(xi - ref)) <= (1e-9 if method == "fftlog" else 1e-5) * _scale(ref)


@pytest.mark.gpu
def test_device_correlation_taper_and_parameters(cosmo):
    """taper_cl through the C entry, and non-default FFTLog sampling parameters"""
    import ccl_b200 as ccl
    from ccl_b200 import _lib, lowlevel
    s, ell, cl = gaussian_cl()
    theta = np.array([0.05, 0.2, 0.7])
    lim = [2., 10., 500., 900.]
    for method, tol in (("fftlog", 1e-9), ("legendre", 1e-9)):
        ref, st = oracle.correlation(ell, cl, theta, "NN", method, taper_cl_limits=lim)
        xi, st2 = lowlevel.correlation(cosmo._device(), ell, cl, theta, _lib.CORR_GG, ccl.correlations.correlation_methods[method],
                                       taper_cl_limits=lim)
        assert st == 0 and st2 == 0
        assert np.max(np.fabs(xi - ref)) <= tol * _scale(ref)
    ccl.spline_params.N_ELL_CORR = 2047          # odd N: no Nyquist term
    ccl.spline_params.ELL_MIN_CORR = 0.05
    ccl.spline_params.ELL_MAX_CORR = 30000.
    try:
        c2 = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks',
                           matter_power_spectrum='linear')
    finally:
        ccl.spline_params.reload()
    for kind in ("NN", "GG-"):
        ref, st = oracle.correlation(ell, cl, theta, kind, "fftlog", ell_min_corr=0.05, ell_max_corr=30000.,
                                     n_ell_corr=2047)
        xi = ccl.correlation(c2, ell=ell, C_ell=cl, theta=theta, type=kind, method="fftlog")
        assert st == 0 and np.max(np.fabs(xi - ref)) <= 1e-9 * _scale(ref)


# ---- the reference's unit tests (pyccl/tests/test_correlations.py) through the pyccl-compatible API ----
@pytest.mark.gpu
@pytest.mark.parametrize('method', ['bessel', 'legendre', 'fftlog'])
def test_correlation_smoke(cosmo, method):
    import ccl_b200 as ccl
    z = np.linspace(0., 1., 200)
    n = np.ones(z.shape)
    lens = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
    ell = np.logspace(1, 3, 5)
    cl = ccl.angular_cl(cosmo, lens, lens, ell)
    t_arr = np.logspace(-2., np.log10(5.), 5)
    t_lst = [t for t in t_arr]
    for tval in [t_arr, t_lst, 2., 2]:
        corr = ccl.correlation(cosmo, ell=ell, C_ell=cl, theta=tval, type='NN', method=method)
        assert np.all(np.isfinite(corr))
        assert np.shape(corr) == np.shape(tval)


def test_correlation_raises():
    import ccl_b200 as ccl
    c = ccl.CosmologyVanillaLCDM(transfer_function='bbks')
    with pytest.raises(ValueError):
        ccl.correlation(c, ell=[1], C_ell=[1e-3], theta=[1], method='blah')
    with pytest.raises(ValueError):
        ccl.correlation(c, ell=[1], C_ell=[1e-3], theta=[1], type='blah')


def test_correlation_zero():
    import ccl_b200 as ccl
    c = ccl.CosmologyVanillaLCDM(transfer_function='bbks')
    ell = np.arange(2, 100000)
    corr = ccl.correlation(c, ell=ell, C_ell=np.zeros(ell.size), theta=np.logspace(0, 2, 10000))
    assert (corr == np.zeros(10000)).all()


@pytest.mark.gpu
def test_correlation_zero_ends(cosmo):
    """an error instead of a crash (pyccl/tests/test_correlations.py:123-131)"""
    import ccl_b200 as ccl
    ell = np.arange(2, 1001)
    C_ell = np.zeros(ell.size)
    C_ell[500] = 1.0
    with pytest.raises(ccl.CCLError):
        ccl.correlation(cosmo, ell=ell, C_ell=C_ell, theta=np.logspace(0, 2, 20))
