"""CPU pins of the oracle (oracle/limber_oracle.c) -- no GPU needed.

The oracle restates GSL's akima / natural-cspline / bicubic / QAG-GK41 arithmetic and the
reference's Limber logic.  GSL is not installed in this image, so each restated algorithm is
pinned against an independent implementation (scipy, closed forms) and against the reference's
own known-answer tests for this path:
  * benchmarks/test_tracers.py:6-130  power-law tracers/P(k) with closed-form C_ell (5e-3)
  * benchmarks/ccl_test_f2d.c          analytic (k/0.1)^-1 f2d cases incl. extrapolation
  * src/ccl_tracers.c:703-726          f_ell values
"""
import numpy as np
import pytest
from scipy.interpolate import Akima1DInterpolator, CubicSpline

import oracle
from common import chi_max_nokernel, oracle_cosmo


def test_akima_matches_scipy_on_generic_data():
    rng = np.random.default_rng(1)
    x = np.sort(rng.uniform(0, 10, 40))
    y = np.sin(x) + 0.1 * rng.standard_normal(40)
    s = oracle.Akima(x, y)
    ref = Akima1DInterpolator(x, y)
    xe = np.linspace(x[0], x[-1], 777)
    assert np.max(np.abs(s(xe) - ref(xe))) < 5e-14
    # integral over the full range against the piecewise antiderivative
    val, st = s.integ(x[0], x[-1])
    assert st == 0
    assert abs(val - ref.integrate(x[0], x[-1])) < 1e-12


def test_akima_gsl_tie_rule():
    # GSL: NE == 0 -> b = m_i, c = d = 0 (interpolation/akima.c akima_calc); scipy differs here.
    x = np.arange(9.0)
    y = np.array([0, 0, 0, 0, 1, 2, 3, 4, 5.0])  # corner at x = 3
    s = oracle.Akima(x, y)
    # intervals [0,1], [1,2]: all neighbouring slopes equal -> linear pieces through the nodes
    assert s(np.array([0.5, 1.5])) == pytest.approx([0.0, 0.0], abs=0)
    assert s(np.array([6.5])) == pytest.approx([3.5], abs=1e-15)
    v, st = s.integ(0.0, 8.0)
    assert st == 0 and np.isfinite(v)


def test_akima_needs_five_points():
    with pytest.raises(ValueError):
        oracle.Akima(np.arange(4.0), np.arange(4.0))
    # ccl_integ_spline on nk < 5 -> CCL_ERROR_MEMORY
    r, st = oracle.integ_spline(np.arange(4.0), np.ones(4))
    assert st != 0


def test_integ_spline_exact_for_cubic():
    x = np.linspace(-1, 2, 31)
    y = 1 + x - 0.5 * x ** 2
    r, st = oracle.integ_spline(x, y)
    exact = (2 + 2 - 8 / 6.) - (-1 + 0.5 + 1 / 6.)
    assert st == 0 and abs(r - exact) < 1e-12


def test_cspline_matches_scipy_natural():
    rng = np.random.default_rng(2)
    x = np.sort(rng.uniform(-3, 3, 33))
    y = np.cos(2 * x) + x
    s = oracle.CSpline(x, y)
    ref = CubicSpline(x, y, bc_type="natural")
    xe = np.linspace(x[0], x[-1], 501)
    assert np.max(np.abs(s(xe) - ref(xe))) < 2e-13
    assert np.max(np.abs(s(xe, 1) - ref(xe, 1))) < 5e-12
    assert np.max(np.abs(s(xe, 2) - ref(xe, 2))) < 5e-10


def _bicubic_numpy(x, y, z, xe, ye):
    """Independent restatement of gsl interp2d/bicubic.c with scipy natural splines."""
    zx = np.array([CubicSpline(x, z[j], bc_type="natural")(x, 1) for j in range(len(y))])
    zy = np.array([CubicSpline(y, z[:, i], bc_type="natural")(y, 1) for i in range(len(x))]).T
    zxy = np.array([CubicSpline(x, zy[j], bc_type="natural")(x, 1) for j in range(len(y))])
    i = min(max(np.searchsorted(x, xe, side="right") - 1, 0), len(x) - 2)
    j = min(max(np.searchsorted(y, ye, side="right") - 1, 0), len(y) - 2)
    dx, dy = x[i + 1] - x[i], y[j + 1] - y[j]
    t, u = (xe - x[i]) / dx, (ye - y[j]) / dy

    def h(s):
        return np.array([2 * s ** 3 - 3 * s ** 2 + 1, -2 * s ** 3 + 3 * s ** 2, s ** 3 - 2 * s ** 2 + s, s ** 3 - s ** 2])
    ht, hu = h(t), h(u)
    F = np.array([[z[j, i], z[j + 1, i], zy[j, i] * dy, zy[j + 1, i] * dy],
                  [z[j, i + 1], z[j + 1, i + 1], zy[j, i + 1] * dy, zy[j + 1, i + 1] * dy],
                  [zx[j, i] * dx, zx[j + 1, i] * dx, zxy[j, i] * dx * dy, zxy[j + 1, i] * dx * dy],
                  [zx[j, i + 1] * dx, zx[j + 1, i + 1] * dx, zxy[j, i + 1] * dx * dy, zxy[j + 1, i + 1] * dx * dy]])
    return ht @ F @ hu


def test_bicubic_matches_independent_restatement():
    rng = np.random.default_rng(3)
    x = np.sort(rng.uniform(0, 5, 17))
    y = np.sort(rng.uniform(0, 1, 9))
    z = np.sin(x)[None, :] * np.exp(y)[:, None] + 0.05 * rng.standard_normal((9, 17))
    s = oracle.Bicubic(x, y, z)
    for xe, ye in rng.uniform([x[0], y[0]], [x[-1], y[-1]], (60, 2)):
        v, st = s(xe, ye)
        assert st == 0
        assert abs(v - _bicubic_numpy(x, y, z, xe, ye)) < 2e-12
    # nodes are reproduced
    v, _ = s(x[4], y[3])
    assert abs(v - z[3, 4]) < 1e-14
    # outside the grid: GSL_EDOM
    _, st = s(x[-1] + 1, y[0])
    assert st != 0


def test_qag41_known_integrals():
    from scipy.integrate import quad
    # kind 1: |x|^p -- polynomials up to degree 61 are integrated exactly by one GK41 panel
    for p in (0, 1, 7, 30):
        r, e, n, st = oracle.qag41_test(1, float(p), 0.0, 1.5)
        assert st == 0 and n == 1
        assert abs(r - 1.5 ** (p + 1) / (p + 1)) < 1e-13 * max(1, 1.5 ** (p + 1))
    # kind 2: 1/(1 + p x^2) (smooth) against the arctangent
    r, e, n, st = oracle.qag41_test(2, 4.0, -1.0, 2.0)
    assert st == 0 and abs(r - (np.arctan(4.0) + np.arctan(2.0)) / 2.0) < 1e-10
    # kind 0: exp(-x^2) cos(p x), oscillatory: adaptive bisection needed at tight tolerance
    r, e, n, st = oracle.qag41_test(0, 40.0, -3.0, 5.0, epsrel=1e-6)
    ref = quad(lambda x: np.exp(-x * x) * np.cos(40.0 * x), -3.0, 5.0, epsabs=1e-14, epsrel=1e-12, limit=400)[0]
    assert st == 0 and n > 1
    assert abs(r - ref) <= max(e, 1e-6 * abs(ref)) + 1e-13
    # kind 1 with a non-integer power has an end-point singularity in the derivative: bisects
    r, e, n, st = oracle.qag41_test(1, 0.5, 0.0, 1.0, epsrel=1e-8)
    assert st == 0 and n > 1 and abs(r - 2 / 3.) < 1e-8


def test_cquad_known_integrals():
    """gsl_integration_cquad restatement (the reference's fallback after GSL_EROUND, src/ccl_cls.c:245-259)."""
    from scipy.integrate import quad
    # polynomials the 17-point rule already reproduces: the first 33 evaluations settle it
    for p in (0, 1, 7):
        r, e, n, st = oracle.cquad_test(1, float(p), 0.0, 1.5)
        assert st == 0 and n == 33
        assert abs(r - 1.5 ** (p + 1) / (p + 1)) < 1e-12 * max(1, 1.5 ** (p + 1))
    # degree 30: the 17- and 33-point interpolants differ, the interval is refined until epsrel holds
    r, e, n, st = oracle.cquad_test(1, 30.0, 0.0, 1.5)
    assert st == 0 and n > 33 and abs(r / (1.5 ** 31 / 31) - 1) < 1e-4
    # smooth rational function against the arctangent
    r, e, n, st = oracle.cquad_test(2, 4.0, -1.0, 2.0, epsrel=1e-10)
    assert st == 0 and abs(r - (np.arctan(4.0) + np.arctan(2.0)) / 2.0) < 1e-11
    # oscillatory integrand: intervals are split, the error estimate bounds the true error
    r, e, n, st = oracle.cquad_test(0, 40.0, -3.0, 5.0, epsrel=1e-8)
    ref = quad(lambda x: np.exp(-x * x) * np.cos(40.0 * x), -3.0, 5.0, epsabs=1e-15, epsrel=1e-13, limit=400)[0]
    assert st == 0 and n > 33
    assert abs(r - ref) <= max(e, 1e-8 * abs(ref)) + 1e-14
    # end-point singularity in the derivative: sqrt(x) on [0, 1]
    r, e, n, st = oracle.cquad_test(1, 0.5, 0.0, 1.0, epsrel=1e-9)
    assert st == 0 and n > 33 and abs(r - 2 / 3.) < 1e-9
    # the same accuracy class as QAG at the Limber tolerance
    rq = oracle.qag41_test(0, 12.0, -2.0, 3.0, epsrel=1e-4)[0]
    rc = oracle.cquad_test(0, 12.0, -2.0, 3.0, epsrel=1e-4)[0]
    ref = quad(lambda x: np.exp(-x * x) * np.cos(12.0 * x), -2.0, 3.0, epsabs=1e-15, epsrel=1e-13)[0]
    assert abs(rq - ref) < 1e-4 * abs(ref) + 1e-12 and abs(rc - ref) < 1e-4 * abs(ref) + 1e-12
    # bad tolerance: GSL_EBADTOL
    assert oracle.cquad_test(2, 1.0, 0.0, 1.0, epsrel=1e-20)[3] == 13


def test_f_ell_values():
    # ccl_cl_tracer_t_get_f_ell (src/ccl_tracers.c:703-726)
    t0 = oracle.Tracer(der_angles=0)
    t1 = oracle.Tracer(der_angles=1)
    t2 = oracle.Tracer(der_angles=2)
    ell = np.array([1.0, 2.0, 10.0, 11.0, 1000.0, 1001.0])
    assert np.all(t0.f_ell(ell) == 1)
    assert np.allclose(t1.f_ell(ell), ell * (ell + 1), rtol=0)
    f2 = t2.f_ell(ell)
    assert f2[0] == 0
    assert f2[1] == pytest.approx(np.sqrt(4 * 3 * 2 * 1.0), rel=1e-15)
    assert f2[2] == pytest.approx(np.sqrt(12 * 11 * 10 * 9.0), rel=1e-15)
    assert f2[3] == pytest.approx(11.5 ** 2 * (1 - 1.25 / 11.5 ** 2), rel=1e-15)
    assert f2[4] == pytest.approx(1000.5 ** 2 * (1 - 1.25 / 1000.5 ** 2), rel=1e-15)
    assert f2[5] == pytest.approx(1001.5 ** 2, rel=1e-15)


def test_tracer_chi_range_rule():
    # src/ccl_tracers.c:626-666: first/last node whose *neighbour* reaches 5e-4 of the maximum
    chi = np.linspace(0, 100, 101)
    w = np.exp(-0.5 * ((chi - 50) / 5.) ** 2)
    t = oracle.Tracer(kernel=(chi, w))
    thr = 5e-4 * w.max()
    lo = chi[np.where(np.abs(w[1:]) >= thr)[0][0]]
    hi = chi[np.where(np.abs(w[:-1]) >= thr)[0][-1] + 1]
    assert t.chi_min == lo and t.chi_max == hi
    # the end nodes themselves evaluate to the extrapolation constant 0 (src/ccl_f1d.c:137-161)
    w1 = np.ones_like(chi)
    t1 = oracle.Tracer(kernel=(chi, w1))
    assert t1.kernel(np.array([0.0, 100.0, 50.0, -1.0, 101.0])).tolist() == [0.0, 0.0, 1.0, 0.0, 0.0]
    # no kernel: [0, chi(A_SPLINE_MIN)]
    t2 = oracle.Tracer(chi_max_nokernel=4321.0)
    assert t2.chi_min == 0 and t2.chi_max == 4321.0


@pytest.mark.parametrize("is_log", [False, True])
def test_f2d_power_law_and_extrapolation(is_log):
    # benchmarks/ccl_test_f2d.c: f(k, a) = (k/0.1)^-1 * g(a).  With log storage and g = exp(a) the
    # table is linear in (ln k, a), which natural cubic splines reproduce exactly.
    lk = np.log(np.geomspace(1e-3, 10, 64))
    a = np.linspace(0.2, 1.0, 12)
    ga = np.exp(a) if is_log else a
    f = (np.exp(lk)[None, :] / 0.1) ** -1 * ga[:, None]
    tab = np.log(f) if is_log else f
    g = oracle.F2D(a=a, lk=lk, fka=tab, extrap_order_lok=1, extrap_order_hik=1,
                   extrap_linear_growth=oracle.F2D_CONSTGROWTH, is_log=is_log, growth_factor_0=1.0,
                   growth_exponent=0)
    assert g.status == 0
    truth = lambda k, aa: (k / 0.1) ** -1 * (np.exp(aa) if is_log else aa)
    for k, aa in ((0.05, 0.5), (1.0, 0.93), (3e-3, 0.21)):
        v, st = g(np.log(k), aa)
        assert st == 0 and v == pytest.approx(truth(k, aa), rel=1e-12 if is_log else 2e-3)
    if is_log:
        for k in (1e-5, 100.0):   # order-1 extrapolation in ln k is exact for a power law
            v, st = g(np.log(k), 0.5)
            assert st == 0 and v == pytest.approx(truth(k, 0.5), rel=1e-10)
        # below amin with constant growth: value at amin times growth_factor_0^0
        v, st = g(np.log(0.1), 0.05)
        assert st == 0 and v == pytest.approx(truth(0.1, 0.2), rel=1e-12)
    # no-extrapolation policy in a: error below amin (src/ccl_f2d.c:214-231)
    g2 = oracle.F2D(a=a, lk=lk, fka=tab, extrap_linear_growth=oracle.F2D_NOEXTRAPOL, is_log=is_log)
    _, st = g2(np.log(0.1), 0.1)
    assert st != 0


def _prediction(ells, chi_i, chi_f, alpha, beta, gamma, der_bessel, der_angles):
    # benchmarks/test_tracers.py:6-18
    exponent = 2 * gamma - alpha - 2 * beta - 1
    fl = np.ones_like(ells)
    if der_bessel == -1:
        fl *= 1. / (ells + 0.5) ** 4
    if der_angles == 2:
        fl *= (ells + 2.) * (ells + 1.) * ells * (ells - 1)
    elif der_angles == 1:
        fl *= ((ells + 1.) * ells) ** 2
    lfac = fl * (ells + 0.5) ** (alpha + 2 * beta)
    norm = (chi_f ** exponent - chi_i ** exponent) / exponent
    return lfac * norm


POWER_LAW_CASES = [(-2., -1., -1., False, True, False, 0, 0), (-2., 0., -1., False, True, False, 0, 0),
                   (-2., -1., 0., False, True, False, 0, 0), (-2., 0., 0., False, True, False, 0, 0),
                   (0., 0., 0., False, True, False, 0, 0), (-2., -1., -1., True, True, False, 0, 0),
                   (-2., 0., -1., True, True, False, 0, 0), (-2., -1., 0., True, True, False, 0, 0),
                   (-2., 0., 0., True, True, False, 0, 0), (0., 0., 0., True, True, False, 0, 0),
                   (-2., 0., 0., False, False, False, 0, 0), (0., 0., 0., False, False, False, 0, 0),
                   (-2., -1., 0., True, True, True, 0, 0), (-2., -1., 0., False, True, True, 0, 0),
                   (-2., -1., -1., False, True, False, -1, 0), (-2., -1., -1., False, True, False, 0, 1),
                   (-2., -1., -1., False, True, False, 0, 2), (-2., -1., -1., False, True, False, -1, 2)]


def power_law_inputs(bg, alpha, beta, gamma, is_factorizable, w_transfer, mismatch):
    """Flat tables of benchmarks/test_tracers.py:57-120 on the synthetic background."""
    a_of, chi_of = bg["chi_of_a"]
    chi_at = lambda z: float(np.interp(1. / (1 + z), a_of, chi_of))
    chii, chif, chimax = chi_at(0.4), chi_at(0.6), chi_at(0.8)
    chiarr = np.linspace(0.1, chimax, 1024)
    if mismatch:
        chiarr_t = chiarr[(chiarr < 0.9 * chif) & (chiarr > 1.1 * chii)]
    else:
        chiarr_t = chiarr
    a_at = lambda c: np.interp(c, chi_of[::-1], a_of[::-1])
    aarr_t = a_at(chiarr_t)[::-1]
    aarr = a_at(chiarr)[::-1]
    lkarr = np.log(10. ** np.linspace(-6, 3, 512))
    wchi = np.ones_like(chiarr)
    wchi[chiarr < chii] = 0
    wchi[chiarr > chif] = 0
    sub = dict(chi=chiarr, w=wchi)
    if w_transfer:
        ta = (chiarr_t ** gamma)[::-1]
        tk = np.exp(beta * lkarr)
        if is_factorizable:
            sub.update(a=aarr_t, lk=lkarr, fk=tk, fa=ta)
        else:
            sub.update(a=aarr_t, lk=lkarr, fka=ta[:, None] * tk[None, :])
    pk = dict(a=aarr, lk=lkarr, pk=np.ones_like(aarr)[:, None] * np.exp(alpha * lkarr)[None, :])
    return sub, pk, chii, chif


@pytest.mark.parametrize("alpha,beta,gamma,is_factorizable,w_transfer,mismatch,der_bessel,der_angles",
                         POWER_LAW_CASES)
@pytest.mark.parametrize("method", ["qag_quad", "spline"])
def test_power_law_closed_form(synth, method, alpha, beta, gamma, is_factorizable, w_transfer, mismatch,
                               der_bessel, der_angles):
    """The reference's known-answer test for this path (benchmarks/test_tracers.py:34-130, 5e-3).

    The a(chi) table is injected exactly like cosmology_distances_from_input; the kernel's top-hat
    edges are where the closed form is least accurate, hence the reference's 5e-3.
    """
    from common import oracle_tracer
    bg = synth["bg"]
    sub, pk, chii, chif = power_law_inputs(bg, alpha, beta, gamma, is_factorizable, w_transfer, mismatch)
    sub.update(der_bessel=der_bessel, der_angles=der_angles)
    cos = oracle_cosmo(bg)
    tr = oracle_tracer(sub, chi_max_nokernel(bg))
    psp = oracle.F2D(a=pk["a"], lk=pk["lk"], fka=pk["pk"], extrap_order_lok=1, extrap_order_hik=2,
                     extrap_linear_growth=oracle.F2D_CCLGROWTH, is_log=False, growth_factor_0=0.0,
                     growth_exponent=2)
    larr = np.linspace(2, 3000, 100)
    m = oracle.INTEG_QAG if method == "qag_quad" else oracle.INTEG_SPLINE
    cl, st = oracle.angular_cl(cos, [tr], [tr], psp, larr, m)
    assert st == 0
    pred = _prediction(larr, chii, chif, alpha, beta, gamma, der_bessel, der_angles)
    # The reference asserts 5e-3 with its default method (qag_quad).  'spline' samples the top-hat
    # kernel every 0.025 in ln k (~15 nodes across the support), so the edge error is O(1e-2) and
    # ell-independent; it is held to 5e-2 here as a sanity bound, not a reference assertion.
    assert np.all(np.abs(cl / pred - 1) < (5e-3 if method == "qag_quad" else 5e-2))


def test_spline_and_qag_agree_on_3x2pt(synth):
    from common import oracle_pk, oracle_tracer
    bg, pk, subs, colls = synth["bg"], synth["pk"], synth["subs"], synth["colls"]
    cmax = chi_max_nokernel(bg)
    cos, psp = oracle_cosmo(bg), oracle_pk(pk)
    otr = [oracle_tracer(s, cmax) for s in subs]
    ells = np.array([2.0, 17.0, 150.0, 1200.0, 3000.0])
    for (i, j) in ((0, 0), (1, 8), (6, 13)):
        t1 = otr[colls[i][0]:colls[i][0] + colls[i][1]]
        t2 = otr[colls[j][0]:colls[j][0] + colls[j][1]]
        a, st_a = oracle.angular_cl(cos, t1, t2, psp, ells, oracle.INTEG_SPLINE)
        b, st_b = oracle.angular_cl(cos, t1, t2, psp, ells, oracle.INTEG_QAG)
        c, st_c = oracle.angular_cl(cos, t2, t1, psp, ells, oracle.INTEG_SPLINE)
        assert st_a == st_b == st_c == 0
        assert np.max(np.abs(a / b - 1)) < 2e-3          # two quadratures of the same integrand
        assert np.max(np.abs(a / c - 1)) < 1e-12         # symmetric under swapping the tracers


def test_disjoint_supports_fail_like_reference():
    # nk < 5 -> gsl_spline_alloc(akima) fails -> NaN + CCL_ERROR_INTEG
    chi = np.linspace(0, 4000, 401)
    a = 1.0 / (1 + chi / 3000.)
    cos = oracle.Cosmo(chi, a)
    w1 = np.where((chi > 1000) & (chi < 1040), 1.0, 0.0)
    w2 = np.where((chi > 1020) & (chi < 1060), 1.0, 0.0)
    t1, t2 = oracle.Tracer(kernel=(chi, w1)), oracle.Tracer(kernel=(chi, w2))
    lk = np.log(np.geomspace(1e-4, 1e2, 50))
    aa = np.linspace(0.1, 1, 8)
    psp = oracle.F2D(a=aa, lk=lk, fka=np.ones((8, 50)), is_log=False)
    cl, st = oracle.angular_cl(cos, [t1], [t2], psp, np.array([10.0, 100.0]), oracle.INTEG_SPLINE)
    if oracle.limber_nk(*oracle.k_interval(cos, [t1], [t2], 10.0)) < 5:
        assert st == 1030 and np.all(np.isnan(cl))
