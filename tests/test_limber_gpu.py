"""GPU parity: the CUDA Limber path (through the C ABI) against the CPU oracle on identical
tables.  Tolerances are BASELINE.json's: 1e-10 relative for 'spline', 1e-5 for 'qag_quad'."""
import numpy as np
import pytest

import oracle
from common import chi_max_nokernel, oracle_cosmo, oracle_pk, oracle_tracer

pytestmark = pytest.mark.gpu

SPLINE_RTOL = 1e-10
QAG_RTOL = 1e-5
ELLS = np.unique(np.geomspace(2, 3000, 24).astype(int)).astype(float)


def _rel(cl, ref, scale_ref):
    # C_ell that cross zero (IA x shear) are compared against the curve's own scale
    return np.abs(cl - ref) / np.maximum(np.abs(ref), 1e-6 * np.max(np.abs(scale_ref)))


@pytest.fixture(scope="module")
def setup(synth):
    import ccl_b200 as cb
    bg, pk, subs, colls = synth["bg"], synth["pk"], synth["subs"], synth["colls"]
    cmax = chi_max_nokernel(bg)
    ocos = oracle_cosmo(bg)
    opk = oracle_pk(pk)
    otr = [oracle_tracer(s, cmax) for s in subs]
    cos = cb.LibCosmology()
    cos.set_achi(bg["chi"], bg["a"])
    cos.set_growth(bg["a_growth"], bg["growth"])
    cos.set_chi_max_nokernel(cmax)
    psp = cb.LibF2D(a=pk["a"], lk=pk["lk"], fka=pk["logp"], extrap_order_lok=1, extrap_order_hik=2,
                    is_log=True, growth_exponent=2, growth_factor_0=0.0)
    ltr = [cb.LibTracer(cos, **s) for s in subs]
    return dict(ocos=ocos, opk=opk, otr=otr, cos=cos, psp=psp, ltr=ltr, colls=colls)


def _coll(objs, coll):
    return objs[coll[0]:coll[0] + coll[1]]


PAIRS = [(0, 0), (0, 3), (1, 7), (4, 14), (5, 5), (6, 12), (2, 2), (0, 9)]


@pytest.mark.parametrize("pair", PAIRS)
def test_single_pair_spline(setup, pair):
    import ccl_b200 as cb
    s = setup
    c1, c2 = s["colls"][pair[0]], s["colls"][pair[1]]
    ref, st_ref = oracle.angular_cl(s["ocos"], _coll(s["otr"], c1), _coll(s["otr"], c2), s["opk"], ELLS,
                                    oracle.INTEG_SPLINE)
    cl, st = cb.angular_cls_limber(s["cos"], _coll(s["ltr"], c1), _coll(s["ltr"], c2), s["psp"], ELLS,
                                   cb._lib.INTEGRATION_SPLINE)
    assert st == st_ref == 0
    assert np.all(np.isfinite(cl))
    assert np.max(_rel(cl, ref, ref)) < SPLINE_RTOL


@pytest.mark.parametrize("pair", PAIRS[:5])
def test_single_pair_qag(setup, pair):
    import ccl_b200 as cb
    s = setup
    c1, c2 = s["colls"][pair[0]], s["colls"][pair[1]]
    ref, st_ref = oracle.angular_cl(s["ocos"], _coll(s["otr"], c1), _coll(s["otr"], c2), s["opk"], ELLS,
                                    oracle.INTEG_QAG)
    cl, st = cb.angular_cls_limber(s["cos"], _coll(s["ltr"], c1), _coll(s["ltr"], c2), s["psp"], ELLS,
                                   cb._lib.INTEGRATION_QAG_QUAD)
    assert st == st_ref == 0
    assert np.max(_rel(cl, ref, ref)) < QAG_RTOL


def test_accessors_match_oracle(setup):
    s = setup
    chi = np.linspace(0.0, 9000.0, 301)
    a_gpu, st = s["cos"].scale_factor_of_chi(chi)
    a_ref, st_ref = s["ocos"].scale_factor_of_chi(chi)
    assert st == st_ref == 0
    assert np.max(np.abs(a_gpu - a_ref)) < 1e-14
    for it in (0, 1, 2, 15, 16):
        w = s["ltr"][it].kernel(chi)
        wr = s["otr"][it].kernel(chi)
        assert np.allclose(w, wr, rtol=1e-13, atol=1e-300)
        assert s["ltr"][it].chi_min == s["otr"][it].chi_min
        assert s["ltr"][it].chi_max == s["otr"][it].chi_max
    lk = np.linspace(np.log(1e-5), np.log(3e3), 97)  # beyond both table ends: extrapolation orders 1 / 2
    for a in (1.0, 0.5, 0.0123, 0.004):              # below amin: growth extrapolation
        p, st = s["psp"].eval(lk, a, s["cos"])
        pr = np.array([s["opk"](x, a, s["ocos"])[0] for x in lk])
        assert st == 0
        inside = (lk >= s["opk"]._lk[0]) & (lk <= s["opk"]._lk[-1])
        assert np.max(np.abs(p / pr - 1)[inside]) < 1e-13
        # Outside the table the order-2 term multiplies d2/dlnk2 of a *natural* spline at its end
        # node, i.e. rounding noise of order eps*|z|/dx^2 ~ 1e-10 in both implementations.
        assert np.max(np.abs(p / pr - 1)[~inside]) < 5e-9
