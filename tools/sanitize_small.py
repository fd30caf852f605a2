"""Small batched problem for compute-sanitizer runs (memcheck / racecheck / synccheck)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccl_b200 as cb  # noqa: E402
from ccl_b200 import synthetic as S  # noqa: E402

bg = S.background()
pk = S.power(bg)
subs, colls = S.lsst_like_tracers(bg, with_ia=True, with_rsd=True, with_mag=True)
a, chi = bg["chi_of_a"]
cmax = float(np.interp(0.1, a, chi))
ells = np.unique(np.geomspace(2, 3000, 10).astype(int)).astype(float)
pairs = S.all_pairs(len(colls))
batch = cb.LibBatch(2)
for ic in range(2):
    batch.set_cosmology(ic, bg["chi"], bg["a"], bg["a_growth"], bg["growth"], cmax)
    batch.set_pk2d(ic, pk["lk"], pk["a"], pk["logp"] + 0.01 * ic)
    for s in subs:
        batch.add_tracer(ic, **s)
batch.commit()
cl, stat, st = batch.angular_cl(colls, pairs, ells)
print("fast", batch.last_used_fast_path, "status", st, "finite", bool(np.all(np.isfinite(cl))))
clq, statq, stq = batch.angular_cl(colls, pairs[:8], ells[:4], cb._lib.INTEGRATION_QAG_QUAD)
print("qag status", stq)
