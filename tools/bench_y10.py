#!/usr/bin/env python
"""BASELINE configs[4]: LSST-Y10 cosmic shear, 20 source bins with IA (210 pairs) x 1000 ells, ONE cosmology.
Times the batched entry on one GPU (the multi-GPU version shards the ell axis: every (pair, ell) is independent)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ccl_b200 as cb  # noqa: E402
from ccl_b200 import synthetic as S  # noqa: E402

bg = S.background()
pk = S.power(bg)
subs, colls = S.lsst_like_tracers(bg, n_lens=0, n_src=20, with_ia=True)
a, chi = bg["chi_of_a"]
cmax = float(np.interp(0.1, a, chi))
ells = np.geomspace(2, 5000, 1000)
pairs = S.all_pairs(len(colls))
batch = cb.LibBatch(1)
batch.set_cosmology(0, bg["chi"], bg["a"], bg["a_growth"], bg["growth"], cmax)
batch.set_pk2d(0, pk["lk"], pk["a"], pk["logp"])
for s in subs:
    batch.add_tracer(0, **s)
batch.commit()
for rep in range(3):
    t0 = time.perf_counter()
    cl, stat, st = batch.angular_cl(colls, pairs, ells)
    dt = time.perf_counter() - t0
print("Y10 shear+IA: %d pairs x %d ells, fast path %s, status %d: %.2f ms per call -> %.3g C_ell/s"
      % (len(pairs), len(ells), batch.last_used_fast_path, st, dt * 1e3, len(pairs) * len(ells) / dt))
if "--check" in sys.argv:
    import oracle
    from common import edge_mask, oracle_cosmo, oracle_pk, oracle_tracer, parity_error
    oc, op = oracle_cosmo(bg), oracle_pk(pk)
    otr = [oracle_tracer(s, cmax) for s in subs]
    worst = 0.0
    for q in range(0, len(pairs), 23):
        i, j = pairs[q]
        t1 = otr[colls[i][0]:colls[i][0] + colls[i][1]]
        t2 = otr[colls[j][0]:colls[j][0] + colls[j][1]]
        redo = lambda: oracle.angular_cl(oc, t1, t2, op, ells[::25], oracle.INTEG_SPLINE)[0]
        worst = max(worst, parity_error(cl[0, q][::25], redo(), redo, 1e-10, edge_mask(t1, t2, ells[::25])))
    print("max rel err vs oracle on a sample: %.3e" % worst)
