#!/usr/bin/env python
"""Stage timings of the parameters -> tables -> batch path (device_builders.build_3x2pt_batch), per step.
usage (GPU box): python tools/fp_profile.py [ncosmo]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccl_b200 as cb  # noqa: E402
from ccl_b200 import device_builders as db, synthetic as S  # noqa: E402
from ccl_b200.cosmology import Cosmology  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
rng = np.random.default_rng(0)
u = rng.random((n, 7))
cos = [dict(Cosmology(Omega_c=0.2 + 0.15 * r[0], Omega_b=0.04 + 0.015 * r[1], h=0.6 + 0.2 * r[2], n_s=0.92 + 0.08 * r[3],
                      sigma8=0.7 + 0.2 * r[4], w0=-1.3 + 0.6 * r[5], wa=-0.5 + r[6],
                      transfer_function='eisenstein_hu')._params) for r in u]
z = np.linspace(0.0, 3.5, 500)
lens = S.smail_bins(z, 0.26, 0.94, edges=np.arange(0.2, 1.2 + 1e-9, 0.2), sigma_z=0.03)
src = S.smail_bins(z, 0.13, 0.78, nbins=10, sigma_z=0.05)
batch = cb.LibBatch(n)
for rep in range(3):
    t0 = time.perf_counter()
    tabs = db.DeviceTables(cos, z, list(lens) + list(src), lens_scale=[0.0] * 5 + [1.0] * 10)
    t1 = time.perf_counter()
    batch.reset()
    tabs.fill(batch)
    t2 = time.perf_counter()
    a_rev = np.ascontiguousarray(1. / (1 + z[::-1]))
    for i in range(5):
        tabs.add_tracer(batch, "nc", i, 0, 0, transfer_a=(a_rev, np.ones(z.size)))
    for j in range(10):
        tabs.add_tracer(batch, "lensing", 5 + j, -1, 2)
    t3 = time.perf_counter()
    batch.commit()
    t4 = time.perf_counter()
    print("n=%d  tables_build %.1f ms | fill %.1f | add_tracer x15 %.1f | commit %.1f | total %.1f" %
          (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3, (t4 - t0) * 1e3))
    del tabs
