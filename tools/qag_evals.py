import sys, numpy as np
sys.path.insert(0,'/root/repo')
import bench, ccl_b200 as cb
tables = bench.build_host_tables(2)
colls, pairs, ells = bench.geometry()
b = cb.LibBatch(2); b.set_stacked(0, bench.stack_host_tables(tables, 2, 0)); b.commit()
cl, stat, st = b.angular_cl(colls, pairs, ells, cb._lib.INTEGRATION_QAG_QUAD)
n = cl.size
ev = b.last_qag_evals if not callable(b.last_qag_evals) else b.last_qag_evals()
print("C_ell", n, "evals", ev, "per C_ell", ev / n, "panels41 per C_ell", ev / n / 41)
# per pair type
for name, sel in (("NCxNC", [q for q,(i,j) in enumerate(pairs) if i<5 and j<5]), ("NCxWL", [q for q,(i,j) in enumerate(pairs) if i<5 and j>=5]), ("WLxWL", [q for q,(i,j) in enumerate(pairs) if i>=5])):
    cl2, _, _ = b.angular_cl(colls, [pairs[q] for q in sel], ells, cb._lib.INTEGRATION_QAG_QUAD)
    ev = b.last_qag_evals if not callable(b.last_qag_evals) else b.last_qag_evals()
    print(name, len(sel), "evals per C_ell %.1f" % (ev / cl2.size))
