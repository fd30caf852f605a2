"""Oracle + host table builders against the reference's own golden C_ell curves
(benchmarks/test_cells.py: 23 tracer-pair types x 541 ells x 2 n(z) families, asserted at
0.1 sigma_CV for qag_quad and 0.2 sigma_CV for spline).  CPU only."""
import numpy as np
import pytest

import oracle
from golden_setup import CASES, SPLINE_CASES, oracle_objects, set_up, sigma_cv


@pytest.fixture(scope="module", params=["analytic", "histo"])
def golden(request):
    cosmo, trc, lfc, bmk = set_up(request.param)
    oc, op, otr = oracle_objects(cosmo, trc)
    return dict(oc=oc, op=op, otr=otr, lfc=lfc, bmk=bmk)


@pytest.mark.parametrize("t1,t2,bm,a1b1,a1b2,a2b1,a2b2,fl", CASES)
def test_cells_qag(golden, t1, t2, bm, a1b1, a1b2, a2b1, a2b2, fl):
    g = golden
    cl, st = oracle.angular_cl(g["oc"], g["otr"][t1], g["otr"][t2], g["op"], g["lfc"]['ells'], oracle.INTEG_QAG)
    assert st == 0
    el = sigma_cv(g["bmk"], g["lfc"], a1b1, a1b2, a2b1, a2b2)
    assert np.all(np.fabs(cl * g["lfc"][fl] - g["bmk"][bm]) < 0.1 * el)


@pytest.mark.parametrize("t1,t2,bm,a1b1,a1b2,a2b1,a2b2,fl", SPLINE_CASES)
def test_cells_spline(golden, t1, t2, bm, a1b1, a1b2, a2b1, a2b2, fl):
    g = golden
    cl, st = oracle.angular_cl(g["oc"], g["otr"][t1], g["otr"][t2], g["op"], g["lfc"]['ells'], oracle.INTEG_SPLINE)
    assert st == 0
    el = sigma_cv(g["bmk"], g["lfc"], a1b1, a1b2, a2b1, a2b2)
    assert np.all(np.fabs(cl * g["lfc"][fl] - g["bmk"][bm]) < 0.2 * el)
