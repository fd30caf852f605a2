"""Diagnostic: one (cosmology, pair, ell) through the fast kernel, the general kernel and the oracle."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ccl_b200 as cb
from ccl_b200 import synthetic as S
import oracle
from common import chi_max_nokernel, oracle_cosmo, oracle_pk, oracle_tracer

ELLS = np.unique(np.geomspace(2, 3000, 24).astype(int)).astype(float)
bg = S.background(Om=0.26, h=0.64)
pk = S.power(bg)
pk["logp"] = pk["logp"] + 0.01
subs, colls = S.lsst_like_tracers(bg, with_ia=True, with_mag=True)
cmax = chi_max_nokernel(bg)
pairs = [(i, j) for i in range(len(colls)) for j in range(i, len(colls))]
batch = cb.LibBatch(1)
batch.set_cosmology(0, bg["chi"], bg["a"], bg["a_growth"], bg["growth"], cmax)
batch.set_pk2d(0, pk["lk"], pk["a"], pk["logp"])
for s in subs:
    batch.add_tracer(0, **s)
batch.commit()
fast, _, _ = batch.angular_cl(colls, pairs, ELLS)
print("fast path:", batch.last_used_fast_path)
os.environ["CCL_B200_NO_FAST"] = "1"
gen, _, _ = batch.angular_cl(colls, pairs, ELLS)
oc, op = oracle_cosmo(bg), oracle_pk(pk)
otr = [oracle_tracer(s, cmax) for s in subs]
ref = np.empty_like(fast[0])
for q, (i, j) in enumerate(pairs):
    ref[q], _ = oracle.angular_cl(oc, otr[colls[i][0]:colls[i][0] + colls[i][1]],
                                  otr[colls[j][0]:colls[j][0] + colls[j][1]], op, ELLS, oracle.INTEG_SPLINE)
sc = np.max(np.abs(ref), axis=1, keepdims=True)
for name, x in (("fast-gen", (fast[0] - gen[0])), ("fast-ref", fast[0] - ref), ("gen-ref", gen[0] - ref)):
    e = np.abs(x) / np.maximum(np.abs(ref), 1e-6 * sc)
    idx = np.argwhere(e > 1e-12)
    print(name, "max", e.max(), "n>1e-12:", len(idx), [(pairs[a], ELLS[b], e[a, b]) for a, b in idx[:10]])
q = pairs.index((14, 14))
l = ELLS[4]
lo, hi = oracle.k_interval(oc, otr[colls[14][0]:colls[14][0] + 2], otr[colls[14][0]:colls[14][0] + 2], l)
x = (hi - lo) / 0.025 + 0.5
print("ell", l, "lkmin", repr(lo), "lkmax", repr(hi), "x", repr(x), "nk", oracle.limber_nk(lo, hi))
