"""The reference's own unit tests of the path (pyccl/tests/test_cells.py, Limber cases), run against
the pyccl-compatible API of this package on the GPU; every curve is also checked against the oracle.

  test_cells_smoke               pyccl/tests/test_cells.py:25-81   six tracer flavours x four ell forms x
                                                                    three p_of_k_a forms, symmetry, invalid dndz
  test_cells_raise_ell_reversed  :84-87
  test_cells_raise_integ_method  :90-101
  test_cells_raise_nonlimber     :104-120 (Limber-only build: unknown method names still raise)
  test_cells_raise_weird_pk      :123-126
  test_cells_calculator          :380-409 without modified gravity: a Calculator fed the P(k,a) arrays of
                                  another cosmology returns the same C_ell to 1e-10
"""
import numpy as np
import pytest

import oracle
from common import edge_mask, parity_error
from golden_setup import oracle_objects

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import ccl_b200 as ccl
    cosmo = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96,
                          transfer_function="bbks", matter_power_spectrum="linear")
    zz = np.linspace(0.0, 1.0, 200)
    nn = np.exp(-(((zz - 0.5) / 0.1) ** 2))
    lens = ccl.WeakLensingTracer(cosmo, dndz=(zz, nn))
    pka = ccl.Pk2D.from_function(pkfunc=lambda k, a: np.log(a / k))
    return dict(ccl=ccl, cosmo=cosmo, zz=zz, nn=nn, lens=lens, pka=pka)


@pytest.mark.parametrize("p_of_k_a", ["default", "pk2d", None])
def test_cells_smoke(env, p_of_k_a):
    ccl, cosmo = env["ccl"], env["cosmo"]
    pk = {"default": ccl.DEFAULT_POWER_SPECTRUM, "pk2d": env["pka"], None: None}[p_of_k_a]
    z = np.linspace(0.0, 1.0, 200)
    n = np.exp(-(((z - 0.5) / 0.1) ** 2))
    b = np.sqrt(1.0 + z)
    lens1 = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
    lens2 = ccl.WeakLensingTracer(cosmo, dndz=(z, n), ia_bias=(z, n))
    nc1 = ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=(z, n), bias=(z, b))
    nc2 = ccl.NumberCountsTracer(cosmo, has_rsd=True, dndz=(z, n), bias=(z, b))
    nc3 = ccl.NumberCountsTracer(cosmo, has_rsd=True, dndz=(z, n), bias=(z, b), mag_bias=(z, b))
    cmbl = ccl.CMBLensingTracer(cosmo, z_source=1100.0)
    tracers = [lens1, lens2, nc1, nc2, nc3, cmbl]
    ells = [4, 4.0, [2, 3, 4, 5], np.arange(2, 5)]
    kw = {} if pk is None else {"p_of_k_a": pk}
    # the oracle on the same host tables (the explicit Pk2D has its own table)
    pk_obj = env["pka"] if p_of_k_a == "pk2d" else None
    oc, op, otr = oracle_objects(cosmo, dict(enumerate(tracers)), pk=pk_obj)
    ell_chk = np.array([2., 3., 4., 5.])
    for i in range(len(tracers)):
        for j in range(i, len(tracers)):
            for ell in ells:
                corr = ccl.angular_cl(cosmo, tracers[i], tracers[j], ell, **kw)
                assert np.all(np.isfinite(corr))
                assert np.shape(corr) == np.shape(ell)
                corr_rev = ccl.angular_cl(cosmo, tracers[j], tracers[i], ell, **kw)
                assert np.allclose(corr, corr_rev)
            # default method (qag_quad) and 'spline' against the oracle
            cq = ccl.angular_cl(cosmo, tracers[i], tracers[j], ell_chk, **kw)
            rq = oracle.angular_cl(oc, otr[i], otr[j], op, ell_chk, oracle.INTEG_QAG)[0]
            assert parity_error(cq, rq, None, 1e-5) < 1e-5, (i, j)
            cs = ccl.angular_cl(cosmo, tracers[i], tracers[j], ell_chk, limber_integration_method="spline", **kw)
            redo = lambda: oracle.angular_cl(oc, otr[i], otr[j], op, ell_chk, oracle.INTEG_SPLINE)[0]  # noqa: E731
            assert parity_error(cs, redo(), redo, 1e-10, edge_mask(otr[i], otr[j], ell_chk)) < 1e-10, (i, j)
    # invalid dndz
    for bad in (z, (z, n, n), (z,), (1, 2)):
        with pytest.raises(ValueError):
            ccl.NumberCountsTracer(cosmo, has_rsd=False, dndz=bad, bias=(z, b))
        with pytest.raises(ValueError):
            ccl.WeakLensingTracer(cosmo, dndz=bad)


@pytest.mark.parametrize("ells", [[3, 2, 1], [1, 3, 2], [2, 3, 1]])
def test_cells_raise_ell_reversed(env, ells):
    with pytest.raises(ValueError):
        env["ccl"].angular_cl(env["cosmo"], env["lens"], env["lens"], ells)


def test_cells_raise_integ_method(env):
    ccl, cosmo, lens = env["ccl"], env["cosmo"], env["lens"]
    with pytest.raises(ValueError):
        ccl.angular_cl(cosmo, lens, lens, [10, 11], limber_integration_method="guad")
    with pytest.raises(ValueError):
        lens2 = ccl.WeakLensingTracer(cosmo, dndz=(env["zz"], np.zeros(len(env["zz"]))))
        ccl.angular_cl(cosmo, lens, lens2, [10, 11], limber_integration_method="quad")


def test_cells_raise_nonlimber(env):
    ccl, cosmo, lens = env["ccl"], env["cosmo"], env["lens"]
    with pytest.raises(ValueError):
        ccl.angular_cl(cosmo, lens, lens, [10, 11], non_limber_integration_method="FEKM")
    with pytest.raises(ValueError):
        ccl.angular_cl(cosmo, lens, lens, [10, 11], l_limber="auoto", non_limber_integration_method="FKEM")


def test_cells_raise_weird_pk(env):
    with pytest.raises(ValueError):
        env["ccl"].angular_cl(env["cosmo"], env["lens"], env["lens"], [10, 11], p_of_k_a=lambda k, a: 10)


def test_cells_calculator(env):
    ccl = env["ccl"]
    cosmo = ccl.CosmologyVanillaLCDM(transfer_function="bbks", matter_power_spectrum="linear")
    cosmo.compute_nonlin_power()
    a, lk, pk = cosmo.get_nonlin_power().get_spline_arrays()
    calc = ccl.CosmologyCalculator(Omega_c=0.25, Omega_b=0.05, h=0.67, n_s=0.96, sigma8=0.81,
                                   pk_nonlin={"a": a, "k": np.exp(lk), "delta_matter:delta_matter": pk})
    ell = np.geomspace(2, 2000, 128)
    tr0 = ccl.CMBLensingTracer(cosmo, z_source=1100.0)
    tr1 = ccl.CMBLensingTracer(calc, z_source=1100.0)
    cl0 = ccl.angular_cl(cosmo, tr0, tr0, ell)
    calc.compute_growth()
    cl1 = ccl.angular_cl(calc, tr1, tr1, ell)
    assert np.all(np.fabs(1 - cl1 / cl0) < 1e-10)


def test_limber_gauss_kronrod_rule_is_honoured_or_rejected():
    """gsl_params.INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS is the GSL rule key gsl_integration_qag receives
    (src/ccl_cls.c:240).  Only GK41 (key 4) exists on the device: any other key must raise for
    'qag_quad' instead of silently integrating with GK41; 'spline' does not read it."""
    import ccl_b200 as ccl
    z = np.linspace(0., 1., 200)
    n = np.ones_like(z)
    ell = np.array([10., 100.])
    try:
        ccl.gsl_params.INTEGRATION_LIMBER_GAUSS_KRONROD_POINTS = 6      # GSL_INTEG_GAUSS61
        cosmo = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks')
    finally:
        ccl.gsl_params.reload()
    t = ccl.WeakLensingTracer(cosmo, dndz=(z, n))
    with pytest.raises(ccl.CCLError, match="GAUSS_KRONROD"):
        ccl.angular_cl(cosmo, t, t, ell, limber_integration_method='qag_quad')
    cs = ccl.angular_cl(cosmo, t, t, ell, limber_integration_method='spline')
    assert np.all(np.isfinite(cs)) and np.all(cs > 0)
    # a cosmology created with the default key integrates
    cosmo2 = ccl.Cosmology(Omega_c=0.27, Omega_b=0.045, h=0.67, sigma8=0.8, n_s=0.96, transfer_function='bbks')
    t2 = ccl.WeakLensingTracer(cosmo2, dndz=(z, n))
    cq = ccl.angular_cl(cosmo2, t2, t2, ell, limber_integration_method='qag_quad')
    assert np.max(np.abs(cq / cs - 1)) < 1e-3


def test_distances_not_computed_is_error_1059():
    """src/ccl_cls.c:274-280: a cosmology without the a(chi) table fails with CCL_ERROR_DISTANCES_INIT."""
    import ccl_b200 as cb
    from ccl_b200 import synthetic as S
    bg = S.background()
    pk = S.power(bg)
    subs, colls = S.lsst_like_tracers(bg, n_lens=1, n_src=1)
    full = cb.LibCosmology()
    full.set_achi(bg["chi"], bg["a"])
    ltr = [cb.LibTracer(full, **s) for s in subs]
    psp = cb.LibF2D(a=pk["a"], lk=pk["lk"], fka=pk["logp"], extrap_order_lok=1, extrap_order_hik=2, is_log=True,
                    growth_exponent=2, growth_factor_0=0.0)
    empty = cb.LibCosmology()           # no distances injected or computed
    cl, st = cb.angular_cls_limber(empty, ltr[:1], ltr[1:2], psp, np.array([10., 20.]), cb._lib.INTEGRATION_SPLINE)
    assert st == 1059       # the output array is left untouched, like the reference's early return


def test_rsd_second_point_reaches_below_pk_a_range(env):
    """The der_bessel = 2 (RSD) sub-tracer evaluates P(k, a') at chi' = (l + 3/2) / k, up to chi_max (l + 3/2) / (l + 1/2):
    with a P(k, a) table that ends just beyond the tracer's chi_max this reaches below the table's a-range at low ell
    and needs the growth table on the device (src/ccl_f2d.c:351-370).  The call must succeed and match the oracle."""
    ccl, cosmo = env["ccl"], env["cosmo"]
    a_arr = np.linspace(1. / 2.3, 1.0, 24)          # P(k, a) tabulated to z = 1.3
    lk_arr = np.log(np.geomspace(1e-4, 50., 96))
    pk_arr = np.array([cosmo.linear_matter_power(np.exp(lk_arr), a) for a in a_arr])
    pk = ccl.Pk2D(a_arr=a_arr, lk_arr=lk_arr, pk_arr=np.log(pk_arr), is_logp=True)
    z = np.linspace(0.8, 1.25, 64)                   # chi_max inside the table, chi_max * 3.5 / 2.5 outside
    nz = np.exp(-0.5 * ((z - 1.05) / 0.06) ** 2)
    nc = ccl.NumberCountsTracer(cosmo, has_rsd=True, dndz=(z, nz), bias=(z, np.ones_like(z)))
    assert nc.chi_max < cosmo.comoving_radial_distance(a_arr[0]) < nc.chi_max * 3.5 / 2.5
    ells = np.array([2., 3., 5., 10., 40., 200.])
    for method, tol in (("spline", 1e-10), ("qag_quad", 1e-5)):
        cl = ccl.angular_cl(cosmo, nc, nc, ells, p_of_k_a=pk, limber_integration_method=method)
        assert np.all(np.isfinite(cl))
        oc, op, otr = oracle_objects(cosmo, {"nc": nc}, pk=pk)
        ref, st = oracle.angular_cl(oc, otr["nc"], otr["nc"], op, ells,
                                    oracle.INTEG_SPLINE if method == "spline" else oracle.INTEG_QAG)
        assert st == 0
        assert np.max(np.abs(cl / ref - 1)) < tol, (method, np.max(np.abs(cl / ref - 1)))
